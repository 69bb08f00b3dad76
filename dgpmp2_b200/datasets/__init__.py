from .planning_dataset import PlanningDataset
from .synthetic import make_problems
