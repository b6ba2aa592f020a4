"""
oracle/build_ref.py -- TEST INFRASTRUCTURE.  Builds the reference's own native code.

Compiles the UNMODIFIED Cython sources of jonathansick/tess from where they lie under
/root/reference (read-only) into oracle/_ref/ (git-ignored, but it travels to the GPU box):

    /root/reference/tess/lloyd.pyx            -> oracle/_ref/lloyd.*.so
    /root/reference/tess/point_accretion.pyx  -> oracle/_ref/point_accretion.*.so

Recipe: Cython `language_level=2` accepts the py2 `print`/`xrange`
syntax; gcc -O2 with no -march (so no FMA contraction), exactly like the reference's
astropy_helpers build would do.  No reference source is copied into the repository: the
generated .c files live only under oracle/_ref/build/.

Usage:  python oracle/build_ref.py [--force]
Returns 0 and prints the built module paths; exits 0 with a message when /root/reference is
absent (the GPU box uses the prebuilt files).
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TESS_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ("lloyd", "point_accretion")


def built_modules():
    return {m: glob.glob(os.path.join(OUT, m + ".*.so")) for m in MODULES}


def build(force=False):
    have = built_modules()
    if all(have.values()) and not force:
        return have
    if not os.path.isdir(os.path.join(REF, "tess")):
        print("oracle/build_ref: %s absent; using prebuilt oracle/_ref (found: %s)"
              % (REF, {k: bool(v) for k, v in have.items()}))
        return have
    import numpy
    from Cython.Build import cythonize
    from setuptools import Extension
    from setuptools.dist import Distribution
    from setuptools.command.build_ext import build_ext

    os.makedirs(os.path.join(OUT, "build"), exist_ok=True)
    exts = [Extension(m, [os.path.join(REF, "tess", m + ".pyx")],
                      include_dirs=[numpy.get_include()],
                      extra_compile_args=["-O2", "-ffp-contract=off", "-w"])
            for m in MODULES]
    exts = cythonize(exts, language_level=2, build_dir=os.path.join(OUT, "build"),
                     quiet=True, force=force)
    dist = Distribution({"name": "tess_ref_oracle", "ext_modules": exts})
    cmd = build_ext(dist)
    cmd.build_lib = OUT
    cmd.build_temp = os.path.join(OUT, "build")
    cmd.inplace = False
    cmd.ensure_finalized()
    cmd.run()
    return built_modules()


if __name__ == "__main__":
    mods = build(force="--force" in sys.argv)
    for k, v in mods.items():
        print(k, "->", v)
