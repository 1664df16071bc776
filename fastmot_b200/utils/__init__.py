from .profiler import Profiler
from .decoder import ConfigDecoder
