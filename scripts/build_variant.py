"""Build a kernel variant of libtef_b200.so into build_variants/ (git-ignored, travels to the GPU box):
    python scripts/build_variant.py merge_any -DTEF_MERGE_MODE=1
    TEF_B200_LIB=build_variants/libtef_merge_any.so python scripts/kernel_times.py ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taming_event_flow_b200 import _lib  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "build_variants", "libtef_%s.so" % name)
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = ["/usr/local/cuda/bin/nvcc"] + _lib.NVCC_FLAGS + flags + ["-I", os.path.join(ROOT, "include"), "-o", out] + _lib.sources()
subprocess.run(cmd, check=True)
print(out)
