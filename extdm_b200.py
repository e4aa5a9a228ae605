"""Import alias: `import extdm_b200` loads the package directory whose name the task fixes
(140-extdm-distribution-extrapolation-diffusion-model-for-video-prediction_b200), which is not a valid
Python identifier."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "140-extdm-distribution-extrapolation-diffusion-model-for-video-prediction_b200")
_spec = importlib.util.spec_from_file_location("extdm_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["extdm_b200"] = _mod
_spec.loader.exec_module(_mod)
