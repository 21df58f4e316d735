"""Locate an (optional) snapshot of the unmodified reference's runtime for the live-plugin
tests and the `kind: "reference"` CPU baseline (oracle/ref_snapshot.sh). Returns the
environment a SUBPROCESS needs. Test infrastructure (lives under oracle/, never imported by the product)."""
import os


def reference_env(root=None):
    root = root or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root
    for base in (os.environ.get("B200PT_MITSUBA_BUILD", ""), os.path.join(root, "oracle", "_ref", "mitsuba_build"), "/tmp/mi_probe/build"):
        if base and os.path.exists(os.path.join(base, "libmitsuba.so")):
            env = dict(os.environ)
            env["PYTHONPATH"] = os.pathsep.join([os.path.join(base, "python"), root, env.get("PYTHONPATH", "")])
            env["LD_LIBRARY_PATH"] = os.pathsep.join([base, os.path.join(base, "python", "mitsuba"), os.path.join(base, "python", "drjit"), env.get("LD_LIBRARY_PATH", "")])
            shim = os.path.join(root, "oracle", "_ref", "llvm_shim", "libLLVM.so")     # oracle/llvm_shim: llvm_ad_rgb without a system libLLVM
            if os.path.exists(shim) and "DRJIT_LIBLLVM_PATH" not in env:
                env["DRJIT_LIBLLVM_PATH"] = shim
            return env
    return None
