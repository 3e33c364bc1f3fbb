import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


@pytest.fixture(scope="session")
def apertures():
    """The reference's aperture masks as its own loader decodes them (byte * float(1/255))."""
    z = np.load(os.path.join(GOLDEN, "apertures.npz"))
    out = {}
    for k in ("pent_11", "pentbig500_14"):
        out[k] = z[k].astype(np.float32) * np.float32(1.0 / 255.0)
        out[k + "_u8"] = z[k]
        out[k + "_total"] = float(z[k + "_total"])
        out[k + "_bbox"] = tuple(int(v) for v in z[k + "_bbox"])
    return out


@pytest.fixture(scope="session")
def port():
    """The C restatement oracle (oracle/lf_oracle.c), compiled on demand with gcc."""
    from oracle import bindings as ob
    ob.build(("port",))
    return ob.PortOracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref).  Built when /root/reference is present,
    used prebuilt on the GPU box, skipped when neither."""
    from oracle import bindings as ob
    try:
        ob.build(("ref",))
    except Exception:
        pass
    if not os.path.exists(ob.REF_SO):
        pytest.skip("oracle/_ref/libref_oracle.so not built (no /root/reference here)")
    return ob.RefOracle()


@pytest.fixture(scope="session")
def native_lib():
    """liblfb200.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from lens_flare_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        capi.build()
    return capi.lib()


@pytest.fixture(scope="session")
def engine(native_lib):
    """One engine on cuda:0 with the built-in RGB lens.  GPU tests fail loudly (not skip) without
    a device: there is no CPU fallback to fall back to."""
    from lens_flare_b200 import capi
    e = capi.Engine(0)
    e.set_lens(capi.builtin_lens(3))
    yield e
    e.close()


def frame_from_golden(golden, name):
    W, H, ax, ay, ang = golden[name + "_meta"]
    W, H = int(W), int(H)
    img = np.zeros((H * W, 3))
    img[golden[name + "_idx"]] = golden[name + "_val"]
    return img.reshape(H, W, 3), W, H, float(ax), float(ay), float(ang)
