"""Drop-in for ERGM's src/model.py: put this directory first on PYTHONPATH (or copy this file over
src/model.py) and `from model import *` in ERGM's src/main.py (line 22) resolves to the B200-native
implementation.  See INTEGRATION.md."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ergm_b200.model import *  # noqa: F401,F403,E402
from ergm_b200.model import __all__  # noqa: F401,E402
