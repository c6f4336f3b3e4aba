"""Dev tool: decode-only GB/s on the shapes of configs B/C/D and two small batches.
A/B the scheduling with SPE_DECODE_VARIANT (unset: dynamic claims, 7: same shape with a static split)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_decode as bd  # noqa: E402

for shape in ((4096, 11, 64, 64), (16384, 17, 96, 72), (2048, 11, 128, 128), (512, 11, 64, 64), (64, 11, 64, 64)):
    bd.run(*shape)
