from .base import BaseManager
from .reward import RewardManager
from .termination import TerminationManager
from .action import BaseActionManager, PositionActionManager, PositionWithinLimitsActionManager
from .command import CommandManager, VelocityCommandManager
from .contact import ContactManager
from .terrain import TerrainManager
from .entity import EntityManager
from .observation import ObservationManager
from .config import MdpFnClass, ResetMdpFnClass

__all__ = [
    "BaseManager", "RewardManager", "TerminationManager", "CommandManager", "VelocityCommandManager",
    "PositionActionManager", "PositionWithinLimitsActionManager", "ContactManager", "TerrainManager",
    "EntityManager", "ObservationManager", "MdpFnClass", "ResetMdpFnClass",
]
