/*
 * swarm_abi.h - C ABI of the fused e-puck swarm step (libswarmstep.so, sm_100a).
 *
 * Drop-in boundary for ONE hot path of ilgha/SwarmACB-isaaclab: the batched 20-robot
 * `env.step` / `env.reset` / `get_critic_state` of the five SwarmACB missions.  The
 * reference has no FFI of its own (pure Python + eager torch); the entry points below are
 * what a binding for this path replaces.  Paths are relative to the reference root; ENV =
 * source/SwarmACB_isaac/SwarmACB_isaac/tasks/direct/missions/directional_gate/directional_gate_env.py,
 * SENS = .../tasks/direct/epuck/epuck_sensors.py, BEH = .../tasks/direct/epuck/behavior_modules.py.
 *
 *   swarm_step          <- DirectMARLEnv.step hook chain: ENV:756-759 (_pre_physics_step),
 *                          ENV:761-843 (_apply_action x decimation), ENV:1200-1209 (_get_dones),
 *                          ENV:1154-1194 + XOR:126 / HOM:87 / FOR:127 / SHL:157 (_get_rewards),
 *                          ENV:1242-1273 (_reset_idx of timed-out envs), ENV:1118-1148
 *                          (_get_observations) incl. SENS:85-501 and BEH:177-574.
 *   swarm_reset         <- DirectMARLEnv.reset: ENV:1242-1273 over all envs + ENV:1118-1148.
 *   swarm_critic_state  <- ENV:1279-1290 -> SENS:545-586.
 *   swarm_rollout       <- the trainers' inner loop `for _ in range(decision_period): env.step(a)`
 *                          (agents/poca_trainer.py:564-573) with device-resident actions.
 *
 * Conventions: plain pointers and sizes, no torch types.  All `float*`/`int*` members of
 * SwarmState/SwarmNoise/SwarmOut are DEVICE pointers (the library never copies them) except in
 * the swarm_host_* entry points, which take HOST buffers and do the copies themselves.
 * Every call only enqueues work on `stream` (a cudaStream_t passed as void*) and never
 * synchronises, except swarm_host_*.  Return value: 0 ok; <0 bad argument (SWARM_E_*);
 * >0 a cudaError_t.
 */
#ifndef SWARM_ABI_H
#define SWARM_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWARM_ABI_VERSION 2
#define SWARM_N 20            /* robots per environment (CFG:39,85) */
#define SWARM_MAX_SEG 16      /* 12 arena faces + <=4 internal walls */
#define SWARM_MAX_INTERNAL 4

enum { SWARM_DGT = 0, SWARM_XOR = 1, SWARM_HOM = 2, SWARM_FOR = 3, SWARM_SHL = 4 };
enum { SWARM_GATE_NONE = 0, SWARM_GATE_DGT = 1, SWARM_GATE_SHL = 2 };

enum {
  SWARM_E_NULL = -1,      /* required pointer is NULL */
  SWARM_E_PARAM = -2,     /* SwarmParams field out of range */
  SWARM_E_SIZE = -3,      /* E <= 0 or steps <= 0 */
  SWARM_E_VERSION = -4,   /* abi_version mismatch */
  SWARM_E_NODEVICE = -5   /* no CUDA device / kernel image for this device */
};

/* Mission / robot constants.  Filled on the host exactly as the reference derives them
 * (double arithmetic, rounded to float32 where torch would cast the Python scalar). */
typedef struct SwarmParams {
  int32_t abi_version;
  int32_t mission;               /* SWARM_DGT.. */
  int32_t obs_dim;               /* 24 (dandelion/daisy/full_policy_observations) or 4 */
  int32_t discrete_actions;      /* 1: module ids int64 (E,N); 0: wheels float (E,N,2) */
  int32_t decimation;            /* CFG:97 */
  int32_t max_episode_length;    /* ceil(episode_length_s / (dt*decimation)) */
  int32_t solver_iterations;     /* CFG:127 */
  int32_t has_light;             /* CFG:172 */
  int32_t n_segments;            /* 12 + n_internal */
  int32_t n_internal;
  int32_t gate_mode;             /* SWARM_GATE_* (ENV:658 / SHL:124 / none) */
  int32_t spawn_max_attempts;    /* CFG:144 */
  float dt, wheelbase, max_wheel_speed;
  float robot_radius, robot_radius_sq, two_radius;
  float wall_r_eff;              /* ENV:1050-1054 */
  float crossing_clearance;      /* ENV:909-913 */
  float capsule_clearance;       /* ENV:986-990 */
  float prox_range, rab_range, rab_loss_probability, unit_scale;
  float light_threshold, light_intensity, alpha;
  float light_x, light_y;
  float critic_radius;
  float spawn_cx, spawn_cy, spawn_sx, spawn_sy, spawn_circle_radius;
  float prox_threshold;          /* BEH:116 */
  /* sensor geometry, float32 results of torch ops on float32 angles (SENS:75-79) */
  float cos_a[8], sin_a[8], rab_cos[4], rab_sin[4];
  /* ztilde = 1 - 2/(1+exp(n)) for n = 0..19 neighbours, float32 torch results (SENS:425) */
  float ztilde_lut[SWARM_N];
  /* arena faces (ENV:849-872) */
  float face_nx[12], face_ny[12], face_px[12], face_py[12];
  /* all raycast segments: ax, ay, bx, by and bx-ax, by-ay in float32 (SENS:204-213) */
  float seg_ax[SWARM_MAX_SEG], seg_ay[SWARM_MAX_SEG], seg_bx[SWARM_MAX_SEG], seg_by[SWARM_MAX_SEG];
  float seg_sx[SWARM_MAX_SEG], seg_sy[SWARM_MAX_SEG];
  /* internal walls for the swept-crossing / capsule solver (ENV:916-938, 993-1015) */
  float iw_ax[SWARM_MAX_INTERNAL], iw_ay[SWARM_MAX_INTERNAL];
  float iw_tx[SWARM_MAX_INTERNAL], iw_ty[SWARM_MAX_INTERNAL];
  float iw_nx[SWARM_MAX_INTERNAL], iw_ny[SWARM_MAX_INTERNAL];
  float iw_len_sq[SWARM_MAX_INTERNAL];
  /* gate push-out: DGT {hw, gate_south, wall_top}; SHL {left,right,bottom,top, r+t/2,
   * bottom-r, top+r, left-r, right+r} */
  float gate[12];
  /* ground zones: DGT {gate_hw, gate_south, corr_south, corr_hw, north_inradius};
   * XOR/FOR/SHL two circles {c0x,c0y,c1x,c1y,r_sq}; HOM {cx,cy,-,-,r_sq};
   * FOR extra {food_radius, nest_top_y}; SHL extra {left,right,bottom,top} */
  float zone[12];
  /* --- manual_control.py StandaloneDGTEnv compatibility (BASELINE config 1), used by swarm_mc_tick only ---
   * MC:531-553 resolves the arena faces one after another (Gauss-Seidel) with r = robot_radius and
   * face data derived from angles; its 12th face duplicates the west face (mid-angle wraps to pi). */
  float mc_face_nx[12], mc_face_ny[12], mc_face_px[12], mc_face_py[12];
  float mc_spawn_safe;       /* MC:251 inradius - 2*robot_radius */
  float mc_spawn_theta_max;  /* MC:253 pi for Homing, 2*pi otherwise */
  int32_t mc_mode;           /* 1 when the block above is filled (build_mc_params) */
} SwarmParams;

/* Persistent per-environment state.  E environments, N = SWARM_N robots. */
typedef struct SwarmState {
  float* pos;                  /* (E,N,2)  ENV:53 agent_pos */
  float* yaw;                  /* (E,N)    ENV:54 */
  float* prev_ground;          /* (E,N)    ENV:61 */
  float* cached_left;          /* (E,N)    ENV:117 */
  float* cached_right;         /* (E,N)    ENV:118 */
  int32_t* fsm;                /* (E,N)    packed BEH:141-153, see SWARM_FSM_* */
  float* beh_cache;            /* (E,6,N)  prox_value, prox_angle, light_value, light_angle,
                                           rab_attr_x, rab_attr_y of the last observation
                                           (ENV:785-795 _sensor_cache); may be NULL when
                                           discrete_actions == 0 */
  uint8_t* mission_flags;      /* (E,N)    bit0 FOR:36 _has_food, bit1 FOR:37 _prev_in_nest */
  int64_t* episode_length_buf; /* (E)      isaaclab DirectMARLEnv */
  float* episode_group_reward; /* (E)      ENV:65 */
  float* completed_group_reward;          /* (E)      ENV:64 */
  float* completed_terminal_critic_state; /* (E,N,5)  ENV:69 */
  int32_t* scratch;            /* >= 4 zero-initialised ints: [0..2] rotating any-reset flags, slot = step_counter % 3
                                  (maintained by the kernels; see swarm_sync_episode_flags); [3] per-step reset
                                  mask of a fused swarm_rollout */
} SwarmState;

/* fsm word layout (bits): explore_state[0] explore_steps[1:4] explore_dir[4:6]
 *                         photo_avoiding[6] photo_steps[7:10] photo_dir[10:12]
 *                         anti_avoiding[12] anti_steps[13:16] anti_dir[16:18];
 * dir encoding 0 -> 0.0, 1 -> +1.0, 2 -> -1.0.
 * Bits 18..23 are not reference state: three 2-bit random numbers the sensor pass of the previous step drew with
 * its packet-loss Philox blocks; the next dispatch takes the duration of a triggered turn (BEH:302, BEH:386) from
 * them (1 + two bits, slot = module 1 / 4 / 5) unless SwarmNoise.turn_dur injects the draw (ABI v2). */

/* Injected noise (parity mode).  Any NULL member falls back to the counter-based Philox
 * stream keyed by (seed, global env index, step counter). */
typedef struct SwarmNoise {
  const float* rab_u;      /* (E,N,N) uniform draws of SENS:420; packet kept iff u >= p_loss */
  const int32_t* turn_dur; /* (E,N,3) BEH:302 / BEH:386 draws for modules 1, 4, 5 */
  const float* spawn_u;    /* (R,E,N,2) rectangle draws per rejection round, ENV:1223 */
  const float* yaw_u;      /* (E,N) ENV:1260 */
  int32_t spawn_rounds;    /* R */
  uint64_t seed;           /* Philox key */
  uint64_t step_counter;   /* Philox counter high word; caller increments per step */
  int64_t env_offset;      /* global index of env 0 of this shard (multi-GPU invariance) */
  const float* rab_u2;     /* (E,N,N) second packet-loss draw of a manual-control tick (MC:435) */
  const float* mc_spawn_u; /* (E,N,3) radius, angle, yaw draws of MC:252-258 */
  /* ENV:1262 re-solves ALL envs of the batch whenever ANY env resets.  By default "the batch" is what this call sees
   * (its E envs: the kernels keep the flag themselves, see swarm_sync_episode_flags).  A job that shards one batch
   * over several GPUs and whose episode counters are not in lockstep can hand in the job-wide flag instead
   * (ABI v2): any_reset_mode = 1 and bit t of any_reset_bits = "some env of the JOB times out at step t of this
   * call" (bit 0 for swarm_step; swarm_rollout then takes at most 32 steps). */
  int32_t any_reset_mode;
  uint32_t any_reset_bits;
} SwarmNoise;

typedef struct SwarmOut {
  float* obs;       /* (E,N,obs_dim) */
  float* reward;    /* (E)  team reward, same for all 20 agents */
  uint8_t* time_out; /* (E) truncated flag */
  float* critic;    /* (E,N,5) or NULL: get_critic_state() of the state the call leaves behind (ENV:1279-1290 ->
                       SENS:545-586), written by the step kernel's epilogue while the pose is still in registers
                       (ABI v2; swarm_rollout writes it after the last step only) */
} SwarmOut;

/* One env.step for E environments.  `actions`: int64 (E,N) module ids when
 * params->discrete_actions, else float (E,N,2) normalised wheel commands. */
int swarm_step(const SwarmParams* params, const SwarmState* state, const void* actions,
               const SwarmNoise* noise, const SwarmOut* out, int E, void* stream);

/* env.reset(): respawn all E environments and produce the first observation. */
int swarm_reset(const SwarmParams* params, const SwarmState* state, const SwarmNoise* noise,
                const SwarmOut* out, int E, void* stream);

/* The reset path re-solves collisions for ALL envs whenever ANY env of the batch times out (ENV:1262).
 * Each step kernel therefore leaves a flag for the next step (slot (step_counter+1) % 3 of state->scratch).
 * A caller that rewrites episode_length_buf or scratch itself, or restarts step_counter, must call this
 * before the next swarm_step so the flag matches the buffer again (swarm_reset does it implicitly). */
int swarm_sync_episode_flags(const SwarmParams* params, const SwarmState* state, uint64_t next_step_counter,
                             int E, void* stream);

/* get_critic_state(): (E,N,5) = (rho, cos a, sin a, cos b, sin b) as its own launch (SwarmOut.critic is the fused
 * alternative). */
int swarm_critic_state(const SwarmParams* params, const SwarmState* state, float* critic_out,
                       int E, void* stream);

/* `steps` consecutive env.step calls with device-resident actions (steps,E,N[,2]); only the
 * last observation is kept, rewards are accumulated into out->reward, time_out is OR-ed.
 * actions_stride_steps = elements between consecutive steps (0 repeats one action: the trainers'
 * decision-period loop).  Runs as ONE fused launch per <= 32 steps (state in registers between
 * steps; with wheel actions the sensors run only after the last one, module actions need them every
 * step).  The results equal `steps` swarm_step calls bit for bit.
 * Injected noise tensors are rejected (single-step only). */
int swarm_rollout(const SwarmParams* params, const SwarmState* state, const void* actions,
                  int64_t actions_stride_steps, const SwarmNoise* noise, const SwarmOut* out,
                  int E, int steps, void* stream);

/* Host-buffer path (what a non-torch caller binds): copies the action batch host->device, runs
 * one step, copies obs/reward/time_out back, synchronises `stream`.  Batches of >= 4096 envs are
 * processed as up to 8 env chunks whose observation downloads run on a library-owned copy stream,
 * overlapping the next chunk's upload + step (use pinned host memory).  `state` holds DEVICE
 * pointers; actions/obs/reward/time_out are HOST pointers; dev_actions is a device staging buffer
 * of the action batch's size. */
int swarm_host_step(const SwarmParams* params, const SwarmState* state, const void* actions_host,
                    const SwarmNoise* noise, float* obs_host, float* reward_host,
                    uint8_t* time_out_host, void* dev_actions, const SwarmOut* dev_out, int E,
                    void* stream);

/* Destroys the library-owned copy streams / events of swarm_host_step (one set per device, created on first use).
 * Optional: call it before cudaDeviceReset or when unloading the library. */
int swarm_host_release(void);

/* One tick of scripts/manual_control.py's loop (MC:721-757) for E standalone environments:
 *   SWARM_MC_PRE     sensors at the current pose (first RAB draw) + BehaviorModules.dispatch without
 *                    previous wheels (MC:729-749); robot 0 keeps the wheel command given in `wheels`
 *   SWARM_MC_PHYSICS StandaloneDGTEnv.step (MC:355-423): clamp in m/s, integrate, one Gauss-Seidel wall
 *                    pass, one gate pass, one robot pass, mission reward, episode roll-over (MC:753-754)
 *   SWARM_MC_POST    compute_obs_robot0's sensor pass (second RAB draw, MC:425-440) -> out->obs (E,N,24)
 * module_ids: int64 (E,N) (needed with PRE); wheels: float (E,N,2) in m/s (robot 0 only with PRE, all
 * robots without).  params must come from the MC parameter builder (mc_mode == 1). */
enum { SWARM_MC_PRE = 1, SWARM_MC_PHYSICS = 2, SWARM_MC_POST = 4 };
int swarm_mc_tick(const SwarmParams* params, const SwarmState* state, const int64_t* module_ids,
                  const float* wheels, const SwarmNoise* noise, const SwarmOut* out, int flags, int E,
                  void* stream);

/* MC:245-269 StandaloneDGTEnv.reset(): polar spawn, no collision re-solve. */
int swarm_mc_reset(const SwarmParams* params, const SwarmState* state, const SwarmNoise* noise, int E,
                   void* stream);

/* Library / device introspection. */
int swarm_abi_version(void);
int swarm_kernel_launch_count(void);       /* kernels launched by this library so far */
const char* swarm_last_error_string(void);

/* Test hook: evaluates include/swarm_detmath.h on the device, sin/cos of a[i] and atan2(a[i], b[i])
 * (device pointers), so tests can check the device results bit for bit against the host's. */
int swarm_detmath_eval(const float* a, const float* b, float* sin_a, float* cos_a, float* atan2_ab, int n,
                       void* stream);

/* FP32 FMA issue-rate micro-benchmark used for the roofline denominator: runs `iters` dependent
 * FMA chains on every SM and returns achieved TFLOP/s in *tflops (synchronises). */
int swarm_fp32_peak(int iters, float* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SWARM_ABI_H */
