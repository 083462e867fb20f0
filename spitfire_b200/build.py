"""
Builds spitfire_b200/libgriffon_b200.so (the C-ABI library of include/griffon_b200.h) with nvcc for sm_100a.

    python -m spitfire_b200.build [--force] [--fma]

nvcc cross-compiles without a GPU. The library links the CUDA runtime statically and has no torch / Python
dependency. The default build passes --fmad=false: the reference is compiled for baseline x86-64 (no FMA,
setup.py:91), and contraction alone moves cancellation-dominated Jacobian entries by up to 1e-9 relative (measured on
the CPU by recompiling the reference algorithm with -mfma). Without contraction the GPU result
stays within 1e-12 of the reference on every well-conditioned entry; the measured cost on k_jac is < 1 %.
`--fma` builds the contracted variant (libgriffon_b200_fma.so) for comparison.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgriffon_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(os.path.dirname(HERE), 'include', 'griffon_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, parity=True, out=LIB, verbose=True, defines=()):
    if not force and not stale(out):
        return out
    flags = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
             '--shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=default',
             '-Xptxas', '-v', '--expt-relaxed-constexpr']
    tag = ' -O3'
    if parity:
        flags += ['--fmad=false']
        tag += ' --fmad=false'
    flags += ['-DGB_FLAGS="%s"' % tag] + ['-D' + d for d in defines]
    cmd = [NVCC] + flags + ['-o', out] + sources()
    if verbose:
        print(' '.join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = res.stdout
    with open(os.path.join(HERE, 'build.log'), 'w') as f:
        f.write(log)
    if res.returncode != 0:
        print(log)
        raise RuntimeError('nvcc failed')
    if verbose:
        for line in log.splitlines():
            if 'registers' in line or 'error' in line or 'warning' in line.lower() and 'ptxas' not in line:
                print(line)
    return out


if __name__ == '__main__':
    if "--timeline" in sys.argv:  # debug build: per-warp phase timeline of k_jac (tools/timeline.py)
        build(force=True, out=os.path.join(HERE, 'libgriffon_b200_tl.so'), defines=('GB_JAC_TIMELINE',))
    elif '--fma' in sys.argv:
        build(force=True, parity=False, out=os.path.join(HERE, 'libgriffon_b200_fma.so'))
    else:
        build(force='--force' in sys.argv)
