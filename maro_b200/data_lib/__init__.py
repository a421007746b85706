"""MARO binary trace format (``.bin``) — writer and reader.

    from maro_b200.data_lib import BinaryConverter, read_bin

``BinaryConverter`` is the counterpart of ``maro.data_lib.BinaryConverter`` (maro/data_lib/binary_converter.py:68): CSV +
meta yaml -> ``.bin``, byte-identical to the reference's output; ``read_bin`` is the reader the scenario loaders use."""
from ..scenarios.citi_bike.data import read_bin
from .binary_converter import BinaryConverter, BinaryMeta

__all__ = ["BinaryConverter", "BinaryMeta", "read_bin"]
