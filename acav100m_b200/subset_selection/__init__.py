"""Drop-in for the reference's ``subset_selection/code`` operator layer."""
from .measures import get_measure  # noqa: F401
from .pairing import get_cluster_pairing  # noqa: F401
