"""The compiled native integrator plugin (native/b200_path_native.cpp) inside the unmodified reference: optional, needs the
runtime of oracle/build_ref.sh with the plugin built by native/build_shim.sh (both travel in oracle/_ref)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle.ref_env import reference_env


def _run(variant):
    env = reference_env(ROOT)
    if env is None or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mitsuba_build", "plugins", "b200_path_native.so")):
        pytest.skip("no reference runtime with the native plugin (oracle/build_ref.sh, native/build_shim.sh)")
    if variant.startswith("llvm") and "DRJIT_LIBLLVM_PATH" not in env:
        pytest.skip("no LLVM for the reference's JIT variant")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_native_shim_live.py"), variant], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NATIVE_SHIM_OK" in r.stdout, r.stdout[-2500:] + r.stderr[-1500:]


@pytest.mark.parametrize("variant", ["scalar_rgb", "llvm_ad_rgb"])
def test_native_plugin_builds_the_scene_description_of_the_python_extractor(built, variant):
    """PluginManager loads native/b200_path_native.so (init_plugin, object.h:343-347), the class instantiates for the variant,
    and the b200pt_scene_desc it assembles through the C++ API (Mesh::packed_vertices / packed_face, Object::traverse) equals
    the Python extractor's: packed records bit for bit, materials, emitters, sensor. Its real render path reaches
    b200pt_scene_create through dlopen; without a GPU that fails loudly through Mitsuba's exception path."""
    _run(variant)


@pytest.mark.gpu
def test_native_plugin_renders_the_same_frame_as_the_python_plugin(built):
    _run("scalar_rgb")
