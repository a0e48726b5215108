python bench.py --steps 10 --warmup 1500 --no-others --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:swarm_kernel -s 1505 -c 1 -f -o gpurun_out/r02_late python bench.py --steps 10 --warmup 1500 --no-others --no-cpu > gpurun_out/r02_late.log 2>&1
ncu -i gpurun_out/r02_late.ncu-rep --page raw --csv > gpurun_out/r02_late_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_late.ncu-rep --page source --csv > gpurun_out/r02_late_sass.csv 2>/dev/null
rm -f gpurun_out/r02_late.ncu-rep
