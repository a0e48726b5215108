"""Importable alias of the ``swarmacb-isaaclab_b200/`` source directory.

The product directory carries the reference's hyphenated name, which Python cannot import; this
package points its ``__path__`` at that directory and executes its ``__init__`` so
``import swarmacb_isaaclab_b200`` (and ``swarmacb_isaaclab_b200.env`` etc.) resolve there.
"""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "swarmacb-isaaclab_b200")
__path__ = [_SRC]
with open(_os.path.join(_SRC, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_SRC, "__init__.py"), "exec"))
del _f
