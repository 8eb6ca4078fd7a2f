"""Installs the UNMODIFIED reference (smdogroup/eigd) into baseline/_ref for `bench.py --impl reference`.

    python baseline/install_reference.py          (build container only: needs /root/reference)

baseline/_ref is git-ignored (never part of the history) but travels to the GPU box with the gpurun snapshot.
Two steps, both offline:
  1. pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref
     of a /tmp copy of the reference (setup.py writes build/ and egg-info into its source tree, and
     /root/reference is read-only); --no-deps because numpy / scipy are already in the image.
  2. the example drivers the timed path lives in (examples/thermal.py:268-342, 560-623 and the modules they
     import) are not part of the `eigd` wheel: they are copied verbatim to baseline/_ref/examples/.
The package does not import under scipy >= 1.15 as shipped; oracle/ref_loader.py loads it with a compatibility
shim for eigd/arpack.py that leaves eigd/eigenvector_derivatives.py and the examples untouched.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("EIGD_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def install(force=False):
    if not os.path.isfile(os.path.join(SRC, "setup.py")):
        return "reference tree %s not present (GPU box: uses the prebuilt baseline/_ref)" % SRC
    marker = os.path.join(DST, "eigd", "eigenvector_derivatives.py")
    if os.path.isfile(marker) and os.path.isdir(os.path.join(DST, "examples")) and not force:
        return "baseline/_ref already installed"
    shutil.rmtree(DST, ignore_errors=True)
    with tempfile.TemporaryDirectory() as tmp:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(SRC, copy)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DST, copy]
        subprocess.run(cmd, check=True, cwd=tmp)
    shutil.copytree(os.path.join(SRC, "examples"), os.path.join(DST, "examples"))
    return "installed %s -> %s" % (SRC, DST)


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
