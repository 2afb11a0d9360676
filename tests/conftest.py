"""Shared fixtures. `-m "not gpu"` runs here (no GPU); `-m gpu` runs on a B200 box where
/root/reference does not exist: GPU tests use only committed fixtures, oracle/ (the C port,
built from source with gcc) and the prebuilt oracle/_ref/ files that travel with the snapshot."""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_package():
    name = "hana_softwarerenderer_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "hana-softwarerenderer_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def hana():
    return load_package()


@pytest.fixture(scope="session")
def horacle():
    from oracle import horacle as H
    return H


@pytest.fixture(scope="session")
def port(horacle):
    return horacle.Port()


@pytest.fixture(scope="session")
def ctx(hana):
    c = hana.Context(0)
    yield c
    c.close()


ASSET_DIR = os.path.join(ROOT, "oracle", "_ref", "assets")


def bundled_scene(hana, name):
    p = os.path.join(ASSET_DIR, name + ".npz")
    if not os.path.exists(p):
        pytest.skip("bundled scene pack %s missing (oracle/pack_assets.py needs /root/reference)" % p)
    return hana.load_hscene(p)


@pytest.fixture(scope="session")
def african_head(hana):
    return bundled_scene(hana, "african_head")


@pytest.fixture(scope="session")
def diablo(hana):
    return bundled_scene(hana, "diablo3_pose")


@pytest.fixture(scope="session")
def blob(hana):
    return hana.synthetic_scene("blob")


def cleared(W, H, rgba=(0, 0, 0, 1), depth=None):
    color = np.empty((H, W, 4), np.uint8)
    color[:] = np.array(rgba, np.uint8)
    d = np.full((H, W), np.float32(3.4028234663852886e38) if depth is None else depth, np.float32)
    return color, d


def compare_frames(color_a, depth_a, color_b, depth_b, primid_a=None, primid_b=None):
    """The parity metrics of BASELINE.json: coverage/prim-ID mismatches, max |depth diff|, max |colour diff| (RGB)."""
    cov_a = depth_a != np.float32(3.4028234663852886e38)
    cov_b = depth_b != np.float32(3.4028234663852886e38)
    out = {
        "coverage_mismatch": int((cov_a != cov_b).sum()),
        "depth_maxdiff": float(np.abs(depth_a.astype(np.float64) - depth_b.astype(np.float64))[cov_a & cov_b].max())
        if (cov_a & cov_b).any() else 0.0,
        "depth_bits_mismatch": int((depth_a.view(np.uint32) != depth_b.view(np.uint32)).sum()),
        "colour_maxdiff": int(np.abs(color_a[..., :3].astype(int) - color_b[..., :3].astype(int)).max()),
        "colour_mismatch_px": int((color_a[..., :3] != color_b[..., :3]).any(-1).sum()),
    }
    if primid_a is not None and primid_b is not None:
        out["primid_mismatch"] = int((primid_a != primid_b).sum())
    return out
