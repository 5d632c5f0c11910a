from torch.distributions.constraints import *  # noqa: F401,F403
from torch.distributions.constraints import dependent_property, independent, real, real_vector, positive  # noqa: F401
