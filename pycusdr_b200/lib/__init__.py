from .filters import rrcosfilter, gaussianFilter, gmskMod
