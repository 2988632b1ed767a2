"""Drop-in for the reference's ``clustering/code`` operator layer."""
from .sgd_clustering import KMeans  # noqa: F401
