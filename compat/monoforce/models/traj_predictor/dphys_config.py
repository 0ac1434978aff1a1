"""Drop-in for monoforce/src/monoforce/models/traj_predictor/dphys_config.py."""
from monoforce_b200.dphys_config import DPhysConfig, robot_geometry  # noqa: F401
