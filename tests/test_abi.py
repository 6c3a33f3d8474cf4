"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/b200q.h
declares; the header is plain C; ctypes struct layouts match the C ones; compute entry points
fail loudly (no CPU fallback) when no sm_100 device is present."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200q.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200q_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from llmq_b200 import lib
    handle = lib.load()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in b200q.h but not exported by libb200q.so"
        assert s in lib.SIGNATURES, f"{s} has no ctypes signature in llmq_b200/lib.py"
    m = re.search(r"#define B200Q_VERSION (\d+)", open(HEADER).read())
    assert handle.b200q_version() == int(m.group(1))


def test_header_is_plain_c_and_struct_layouts_match(tmp_path):
    from llmq_b200 import lib
    src = tmp_path / "t.c"
    src.write_text(
        '#include "b200q.h"\n#include <stdio.h>\n#include <stddef.h>\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(b200q_model_config), sizeof(b200q_batch),"
        " sizeof(b200q_engine_config), sizeof(b200q_engine_stats), sizeof(b200q_profile),"
        " offsetof(b200q_batch, token_ids), offsetof(b200q_batch, sum_ctx_dec));return 0;}\n")
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(lib.ModelConfig), C.sizeof(lib.Batch), C.sizeof(lib.EngineConfig),
                   C.sizeof(lib.EngineStats), C.sizeof(lib.Profile), lib.Batch.token_ids.offset,
                   lib.Batch.sum_ctx_dec.offset]


def test_no_cpu_fallback():
    import torch
    from llmq_b200 import lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    handle = lib.load()
    assert handle.b200q_device_check() == -5
    with pytest.raises(lib.B200QError):
        lib.require_device()
    cfg = lib.ModelConfig(hidden=256, n_layers=1, n_q_heads=4, n_kv_heads=1, head_dim=64, intermediate=256,
                          vocab=256, block_size=16, max_tokens=16, max_seqs=4, max_pos=64,
                          tie_embeddings=0, rms_eps=1e-5, attn_scale=0.125)
    h = C.c_void_p()
    assert handle.b200q_model_create(C.byref(cfg), C.byref(h)) == -5  # refuses without a device
    # argument validation happens before any CUDA call
    assert handle.b200q_gemm_bf16(None, None, None, 4, 100, 64, None) == -1
    assert b"gemm" in handle.b200q_last_error()


def test_product_never_imports_the_oracle():
    """only tests/, smoke() and bench.py's CPU legs may touch oracle/"""
    pkg = os.path.join(ROOT, "llmq_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
