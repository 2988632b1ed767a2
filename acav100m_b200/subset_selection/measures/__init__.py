"""Measure registry with the reference's plugin point (``measures/__init__.py:5-14``)."""
from .batch_mi import EfficientBatchMI
from .dense_mi import EfficientAMI, EfficientMI
from .mem_mi import EfficientMemMI


def get_measure(measure_name):
    dt = {
        'mi': EfficientMI,
        'ami': EfficientAMI,
        'mem_mi': EfficientMemMI,
        'batch_mi': EfficientBatchMI,
    }
    measure_name = measure_name.lower()
    assert measure_name in dt, "no measure named {}".format(measure_name)
    return dt[measure_name]
