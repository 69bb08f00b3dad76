from .gp_factor import GPFactor
from .prior_factor import PriorFactor
