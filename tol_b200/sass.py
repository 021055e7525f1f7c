"""Identify a kernel of the built library by the hash of its SASS (harness code: bench.py and tools/ use it to tie
an ncu capture under profiles/ to the kernel binary that was captured)."""
import hashlib
import os
import re
import shutil
import subprocess

from .lib import LIB_PATH


def _cuobjdump():
    return shutil.which("cuobjdump") or ("/usr/local/cuda/bin/cuobjdump" if os.path.exists("/usr/local/cuda/bin/cuobjdump") else None)


def bench_kernel_pattern(mission, wind, ts):
    """itanium-mangled template argument list of the plain F+G instance of kernel A the batch path launches for this
    problem: fg_cta_kernel<FORM, WIND, MAXT, MINB, MODE_PLAIN, LOOP = true>"""
    form = 10 if str(mission) in ("S10", "10") else 7
    maxt, minb = (128, 4) if ts <= 128 else (256, 2)
    return "fg_cta_kernelILi%dELi%dELi%dELi%dELi0ELb1EE" % (form, int(wind), maxt, minb)


def kernel_sass_hash(pattern, lib_path=None):
    """first 16 hex digits of the sha256 of the kernel's SASS mnemonics and operands (addresses and encodings
    stripped); None when cuobjdump or the kernel cannot be found"""
    exe = _cuobjdump()
    lib_path = lib_path or LIB_PATH
    if not exe or not os.path.exists(lib_path):
        return None
    try:
        syms = subprocess.run([exe, "-symbols", lib_path], capture_output=True, text=True, timeout=120).stdout
        full = [ln.split()[-1] for ln in syms.splitlines() if pattern in ln and "STO_ENTRY" in ln]
        if len(full) != 1:
            return None
        out = subprocess.run([exe, "-sass", "-fun", full[0], lib_path], capture_output=True, text=True, timeout=120).stdout
    except (OSError, subprocess.TimeoutExpired):
        return None
    body = []
    for ln in out.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?;)", ln)
        if m:
            body.append(m.group(1).strip())
    if len(body) < 100:
        return None
    return hashlib.sha256("\n".join(body).encode()).hexdigest()[:16]
