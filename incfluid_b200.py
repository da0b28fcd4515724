"""Import shim: the package directory is named `incremental-fluids-java_b200` (not a Python
identifier), so `import incfluid_b200` loads it through importlib and re-exports it."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("incremental-fluids-java_b200")
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
package = _pkg
