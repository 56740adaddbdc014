"""Data utilities: API mirror of reference cytofpy/util (simdata.py:11-86,
readCB.py:3-40) plus the GPU-scale synthetic generator the benchmark uses."""
from .synth import simdata, simdata_with_missing_values, synth_cytof, inject_missing
from .cbfile import readCB, preprocess

__all__ = ['simdata', 'simdata_with_missing_values', 'synth_cytof', 'inject_missing',
           'readCB', 'preprocess']
from .hostbind import on_gpu_node, gpu_local_cpus
