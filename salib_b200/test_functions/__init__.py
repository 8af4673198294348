"""Device evaluation of the reference's own test models on the Sobol' path."""
from . import Ishigami, Sobol_G  # noqa: F401
