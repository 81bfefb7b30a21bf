from .vspace_state_var import VSpaceStateVar
from . import vspace_evaluators
