"""Industrial environments (CUDA-resident). Mirrors reference environments/__init__.py."""
from .base import IndustrialEnv
from .chemical_reactor import ChemicalReactorEnv
from .power_grid import PowerGridEnv
from .robot_assembly import RobotAssemblyEnv

__all__ = ["IndustrialEnv", "ChemicalReactorEnv", "PowerGridEnv", "RobotAssemblyEnv"]
