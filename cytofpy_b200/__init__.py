"""cytofpy_b200 — B200-native drop-in for the ADVI training hot path of
luiarthur/CytofPy (`import cytofpy_b200 as cytofpy`)."""
from . import model
from . import util

__version__ = '0.1.0'
