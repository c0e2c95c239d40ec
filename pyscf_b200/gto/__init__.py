from .mole import Mole, M, BOHR
from . import basis
