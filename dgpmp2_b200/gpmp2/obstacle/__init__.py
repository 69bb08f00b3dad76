from .obstacle_factor import ObstacleFactor
from .obstacle_cost import HingeLossObstacleCost
