from .env_2d import Env2D
