"""dev: FP64 pipe peak of the device (DFMA and DMUL+DADD micro-benchmarks of the library)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from spitfire_b200 import griffon
for kind, name in ((0, 'DFMA'), (1, 'DMUL+DADD')):
    tf, ipc = griffon.measure_fp64_peak(kind)
    print(f'{name}: {tf:.2f} Tflop/s, {ipc:.1f} FP64 thread instructions / clk / SM')
for kind, name in ((2, 'DFMA'), (3, 'DADD'), (4, 'DMUL')):
    print(f'{name}: {griffon.measure_fp64_latency(kind):.1f} cycles per dependent instruction (one warp)')
