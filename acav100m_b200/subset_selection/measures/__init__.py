"""Measure registry with the reference's plugin point (``measures/__init__.py:5-14``)."""
from .batch_mi import EfficientBatchMI
from .dense_mi import EfficientMI
from .mem_mi import EfficientMemMI

_PENDING = {
    'ami': "adjusted MI (reference measures/mi.py:212-260)",
}


def get_measure(measure_name):
    name = measure_name.lower()
    if name == 'mem_mi':
        return EfficientMemMI
    if name == 'batch_mi':
        return EfficientBatchMI
    if name == 'mi':
        return EfficientMI
    assert name in _PENDING, "no measure named {}".format(measure_name)
    raise NotImplementedError(
        "measure '{}' -- {} -- is outside the CUDA hot path built so far (DESIGN.md, scope table); "
        "use measure_name='mem_mi'".format(name, _PENDING[name]))
