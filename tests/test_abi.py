"""CPU: the C-ABI library loads, exports every symbol include/loco_asr.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest
import torch

from loco_asr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "loco_asr.h")).read()
    return sorted(set(re.findall(r"LOCO_API\s+[\w\s\*]+?\b(loco_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("loco_create", "loco_destroy", "loco_load_tensor", "loco_finalize_weights", "loco_plan", "loco_encode",
              "loco_encode_host", "loco_last_error", "loco_abi_version"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/loco_asr.h but not exported"
    assert set(declared_symbols()) == set(_lib.SIGNATURES), "ctypes binding and header drifted"
    assert _lib.load().loco_abi_version() == _lib.ABI_VERSION == 2
    for s in ("loco_plan_create", "loco_encode_planned", "loco_plan_destroy", "loco_sync_check"):
        assert s in declared_symbols()


def test_status_and_dtype_codes_match_the_header():
    """The binding's numeric codes are the header's enums (loco_status, loco_dtype, LOCO_POOL_*), and null handles / null
    arguments come back as LOCO_ERR_INVALID without a device (no compute is called here)."""
    text = open(os.path.join(ROOT, "include", "loco_asr.h")).read()
    enums = dict(re.findall(r"\b(LOCO_(?:OK|ERR_\w+|F32|F16|BF16|F64))\s*=\s*(-?\d+)", text))
    for name, val in enums.items():
        assert getattr(_lib, name) == int(val), name
    assert len(enums) == 10
    assert int(re.search(r"#define LOCO_ABI_VERSION (\d+)", text).group(1)) == _lib.ABI_VERSION
    lib = _lib.load()
    ws = C.c_size_t()
    assert lib.loco_finalize_weights(None) == _lib.LOCO_ERR_INVALID
    assert lib.loco_plan(None, None, 0, None, None, None, C.byref(ws)) == _lib.LOCO_ERR_INVALID
    assert lib.loco_encode(None, None, None, 0, None, None, None, 0, None) == _lib.LOCO_ERR_INVALID
    assert lib.loco_encode_text(None, None, None, 0, None, None, None, 0, None) == _lib.LOCO_ERR_INVALID
    assert lib.loco_encode_host(None, None, None, 0, None, None, None, 0, None) == _lib.LOCO_ERR_INVALID
    assert lib.loco_plan_info(None, None, None, None, None) == _lib.LOCO_ERR_INVALID
    assert lib.loco_set_head(None, 0, None, None, None, 0) == _lib.LOCO_ERR_INVALID
    assert lib.loco_sync_check(None, None) == _lib.LOCO_ERR_INVALID
    assert lib.loco_launch_count(None) == 0
    lib.loco_destroy(None)
    lib.loco_plan_destroy(None, None)
    assert lib.loco_create(None, 0, None) == _lib.LOCO_ERR_INVALID and b"null" in lib.loco_last_error(None)


def test_product_library_ships_only_product_kernels():
    """The cross-check kernels (SIMT GEMM, single-CTA tcgen05 GEMM, mma.sync positional conv and attention) live only in the
    LOCO_DEBUG twin; the product library's device code does not contain them, and its attention path has no mma.sync."""
    import subprocess
    assert os.path.exists(_lib.DEBUG_LIB_PATH)
    assert _lib.load().loco_is_debug_build() == 0 and _lib.load(debug=True).loco_is_debug_build() == 1

    def kernels(path):
        out = subprocess.run(["cuobjdump", "-elf", path], capture_output=True, text=True).stdout
        return set(re.findall(r"\.text\.(\w+)", out))

    prod, dbg = kernels(_lib.LIB_PATH), kernels(_lib.DEBUG_LIB_PATH)
    if not prod:
        pytest.skip("cuobjdump not available")
    joined = " ".join(sorted(prod))
    for name in ("gemm_simt", "gemm_tc_kernel", "posconv_kernel", "posconv_tc_kernel", "conv0_mma_kernel", "attention_kernel"):
        assert not re.search(r"\d+%s" % name, joined), name
        assert re.search(r"\d+%s" % name, " ".join(sorted(dbg))), name
    for name in ("gemm_tc2_kernel", "attention_tc_kernel", "posconv_pp_kernel", "conv0_tc_kernel", "final_ln_pool_kernel"):
        assert re.search(name, joined), name


def test_default_config_struct_matches_python_config():
    from loco_asr_b200.config import LocoSpeechT5Config
    lib = _lib.load()
    cc = _lib.LocoConfigC()
    lib.loco_default_config(C.byref(cc))
    cfg = LocoSpeechT5Config()
    assert cc.hidden_size == cfg.hidden_size and cc.encoder_layers == cfg.encoder_layers
    assert list(cc.conv_kernel)[:7] == list(cfg.conv_kernel) and list(cc.conv_stride)[:7] == list(cfg.conv_stride)
    assert cc.encoder_max_relative_position == cfg.encoder_max_relative_position
    assert abs(cc.layer_norm_eps - cfg.layer_norm_eps) < 1e-12


def test_unsupported_config_is_rejected_with_a_message():
    lib = _lib.load()
    cc = _lib.LocoConfigC()
    lib.loco_default_config(C.byref(cc))
    cc.hidden_size = 1024
    h = C.c_void_p()
    assert lib.loco_create(C.byref(cc), 0, C.byref(h)) == -1
    assert b"hidden_size" in lib.loco_last_error(None)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_no_cpu_fallback():
    from loco_asr_b200.encoder import LocoSpeechT5Encoder
    with pytest.raises(_lib.LocoError, match="no CPU fallback"):
        LocoSpeechT5Encoder(device="cuda:0")
    with pytest.raises(_lib.LocoError, match="no CPU fallback"):
        LocoSpeechT5Encoder(device="cpu")
    lib = _lib.load()
    cc = _lib.LocoConfigC()
    lib.loco_default_config(C.byref(cc))
    h = C.c_void_p()
    assert lib.loco_create(C.byref(cc), 0, C.byref(h)) == -2
    assert b"no CPU fallback" in lib.loco_last_error(None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "loco_asr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
