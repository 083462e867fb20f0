"""
build_oracle.py -- TEST INFRASTRUCTURE. Builds the two CPU checkers:

  oracle/liboracle_port.so       plain-C restatement (oracle/griffon_oracle.c), gcc -O2, no FMA contraction
  oracle/_ref/libref_griffon.so  the UNMODIFIED reference Griffon C++ compiled where it lies under
                                 /root/reference/src/spitfire/griffon (reference flags -O3 -g -std=c++11,
                                 setup.py:91) behind oracle/ref_shim.cpp; only built when /root/reference exists.

Outputs go to oracle/ and oracle/_ref/ only (both git-ignored; they travel to the GPU box with gpurun).
The reference's own build system (setup.py / Cython) is not run.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_GRIFFON = '/root/reference/src/spitfire/griffon'


def _scipy_openblas():
    import scipy
    libdir = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), 'scipy.libs')
    libs = sorted(glob.glob(os.path.join(libdir, 'libscipy_openblas*.so')))
    if not libs:
        raise RuntimeError('no libscipy_openblas in ' + libdir)
    return libdir, os.path.basename(libs[0])


def _run(cmd):
    print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_port(force=False):
    src = [os.path.join(HERE, 'griffon_oracle.c'), os.path.join(HERE, 'lapack_shim.c')]
    hdr = [os.path.join(HERE, 'griffon_oracle.h')]
    out = os.path.join(HERE, 'liboracle_port.so')
    if force or _stale(out, src + hdr):
        libdir, lib = _scipy_openblas()
        _run(['gcc', '-O2', '-g', '-std=c11', '-ffp-contract=off', '-fPIC', '-shared', '-Wall', '-Wno-unused-variable',
              '-I' + HERE, '-o', out] + src +
             ['-L' + libdir, '-l:' + lib, '-Wl,-rpath,' + libdir, '-lm'])
    return out


def build_ref(force=False):
    """returns the path, or None when the reference tree is absent and no prebuilt library exists"""
    outdir = os.path.join(HERE, '_ref')
    out = os.path.join(outdir, 'libref_griffon.so')
    if not os.path.isdir(REF_GRIFFON):
        return out if os.path.exists(out) else None
    ref_src = sorted(glob.glob(os.path.join(REF_GRIFFON, 'src', '*.cpp')))
    ref_src = [s for s in ref_src if not s.endswith('flamelet2d_kernels.cpp')] + \
              [s for s in ref_src if s.endswith('flamelet2d_kernels.cpp')]
    mine = [os.path.join(HERE, 'ref_shim.cpp'), os.path.join(HERE, 'lapack_shim.c')]
    if force or _stale(out, ref_src + mine + [os.path.join(HERE, 'griffon_oracle.h')]):
        os.makedirs(outdir, exist_ok=True)
        libdir, lib = _scipy_openblas()
        objs = []
        for s in ref_src + [mine[0]]:
            o = os.path.join(outdir, os.path.basename(s) + '.o')
            _run(['g++', '-O3', '-g', '-std=c++11', '-w', '-fPIC', '-I' + os.path.join(REF_GRIFFON, 'include'),
                  '-I' + HERE, '-c', s, '-o', o])
            objs.append(o)
        o = os.path.join(outdir, 'lapack_shim.o')
        _run(['gcc', '-O2', '-fPIC', '-c', mine[1], '-o', o])
        objs.append(o)
        _run(['g++', '-shared', '-o', out] + objs + ['-L' + libdir, '-l:' + lib, '-Wl,-rpath,' + libdir])
        for o in objs:
            os.remove(o)
    return out


if __name__ == '__main__':
    force = '--force' in sys.argv
    if os.path.exists(os.path.join(HERE, 'griffon_oracle.c')):
        print(build_port(force))
    print(build_ref(force))
