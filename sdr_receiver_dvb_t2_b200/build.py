"""Build libt2b200.so (CUDA kernels + C-ABI) in-tree with nvcc for sm_100a.

Cross-compiles without a GPU.  Rebuilds only when a source is newer than the library.
"""
import fcntl
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libt2b200.so')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def _nvcc():
    for c in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found: libt2b200.so cannot be built (there is no CPU fallback)')


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')) + glob.glob(os.path.join(CSRC, '*.cpp')))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        glob.glob(os.path.join(CSRC, '*.inc')) + [os.path.join(HERE, '..', 'include', 't2b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Several processes may get here at once (one rank per GPU under torchrun): the build runs under a file lock, the
    library appears atomically (linked under a temporary name, then renamed), latecomers find it up to date."""
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    with open(os.path.join(objdir, '.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB
            return _build_locked(objdir, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(objdir, verbose):
    nvcc = _nvcc()
    objs = []
    procs = []
    for s in sources():
        o = os.path.join(objdir, os.path.basename(s) + '.o')
        objs.append(o)
        cmd = [nvcc, '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xptxas', '-v', *ARCH,
               '-I', os.path.join(HERE, '..', 'include'), '-c', s, '-o', o]
        if s.endswith('.cpp'):
            cmd.insert(1, '-x'), cmd.insert(2, 'cu')
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (s, out))
    with open(os.path.join(objdir, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    tmp = LIB + '.tmp.%d' % os.getpid()
    subprocess.run([nvcc, '-shared', *ARCH, '-o', tmp, *objs, '-lcudart'], check=True)
    os.replace(tmp, LIB)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
