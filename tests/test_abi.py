"""The C-ABI library loads and exports every symbol include/solarosc_b200.h declares; the ctypes
mirrors have the C struct sizes.  No compute call is made (no GPU needed)."""
import ctypes as C
import os
import re
import shutil
import subprocess

import pytest

from dune_solaroscfitter_b200 import _lib, abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "solarosc_b200.h")


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.load()


def _declared():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(solarosc_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names          # and the Python loader binds exactly that set


def test_abi_version(lib):
    assert lib.solarosc_abi_version() == abi.ABI_VERSION == 3


def test_struct_layouts_match_the_header(tmp_path):
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n",'
                   'sizeof(solarosc_detector),sizeof(solarosc_tables),sizeof(solarosc_config),'
                   'offsetof(solarosc_tables,det),offsetof(solarosc_detector,obs_cc_night),sizeof(solarosc_detector_ext),'
                   'offsetof(solarosc_detector,ext),offsetof(solarosc_detector_ext,angular_mask));return 0;}\n' % HEADER)
    exe = tmp_path / "sz"
    subprocess.run([cc, str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [C.sizeof(abi.Detector), C.sizeof(abi.Tables), C.sizeof(abi.Config),
                   abi.Tables.det.offset, abi.Detector.obs_cc_night.offset, C.sizeof(abi.DetectorExt),
                   abi.Detector.ext.offset, abi.DetectorExt.angular_mask.offset]


def test_create_rejects_bad_arguments_without_a_gpu(lib, tables):
    st = C.c_int(-1)
    assert not lib.solarosc_create(None, None, C.byref(st))
    assert st.value == abi.ERR_ARG
    packed = abi.PackedTables(tables)
    cfg = abi.make_config()
    cfg.abi_version = 99
    assert not lib.solarosc_create(C.byref(packed.struct), C.byref(cfg), C.byref(st))
    assert st.value == abi.ERR_ARG and b"abi_version" in lib.solarosc_last_error(None)


def test_no_fallback_without_a_device(lib, tables):
    """Without an sm_100 device create must fail loudly with SOLAROSC_ERR_CUDA (never compute on the CPU)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from dune_solaroscfitter_b200.engine import SolarOscEngine, SolarOscError
    with pytest.raises(SolarOscError) as ei:
        SolarOscEngine(tables)
    assert ei.value.status == abi.ERR_CUDA


def test_packed_tables_shapes(tables):
    p = abi.PackedTables(tables)
    assert (p.n_etrue, p.n_cosz, p.n_paths, p.n_layers) == (40, 50, 50, 6)
    assert p.n_params() == 19 and p.n_params(earth_uncertainty=False) == 13
    names = p.param_names(det_names=["HD"])
    assert names[:3] == ["sin2_th12", "sin2_th13", "dm2_21"] and names[8] == "earth_norm_1"
    assert names[-5:] == ["HD_ES_bg_norm_1", "HD_ES_bg_norm_2", "HD_ES_bg_norm_3", "HD_CC_bg_norm_1", "HD_CC_bg_norm_2"]
