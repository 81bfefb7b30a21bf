from .se3_state_var import SE3StateVar
from . import se3_evaluators
