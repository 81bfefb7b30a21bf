from .evaluable import Evaluable, Node, Jacobians
from .state_var import StateVar
