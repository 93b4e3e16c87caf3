"""Compilation of the native pieces: the fixed host runtime
(``libopty_b200.so``) and the per-problem sm_100a modules emitted by
:mod:`opty_b200.codegen`.

Plays the role of the ``setup.py build_ext --inplace`` subprocess and the
source-hash module cache of ``ufuncify_matrix`` (opty/utils.py:759-770,
824-916).  Like the reference, a failed compilation surfaces as an
``ImportError`` carrying the compiler's stderr (opty/utils.py:909-916).
"""

import hashlib
import json
import logging
import os
import shutil
import subprocess
import tempfile

logger = logging.getLogger(__name__)

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
RUNTIME_SRC = os.path.join(CSRC, 'runtime.cu')
RUNTIME_LIB = os.path.join(_HERE, 'libopty_b200.so')
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), 'include')

ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']


def nvcc_path():
    for cand in (os.environ.get('OPTY_B200_NVCC'), shutil.which('nvcc'),
                 '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise ImportError('nvcc was not found; opty_b200 needs the CUDA toolkit '
                      'to build its sm_100a kernels.')


def default_cache_dir():
    return os.environ.get('OPTY_B200_CACHE', os.path.join(_HERE, '_cache'))


def _newer(src_files, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_files)


def build_runtime(force=False, verbose=False):
    """Builds ``libopty_b200.so`` (in-tree) if it is missing or stale."""
    deps = [RUNTIME_SRC, os.path.join(CSRC, 'colloc_params.h'),
            os.path.join(INCLUDE_DIR, 'opty_b200.h')]
    if not force and not _newer(deps, RUNTIME_LIB):
        return RUNTIME_LIB
    tmp_lib = '{}.{}.tmp'.format(RUNTIME_LIB, os.getpid())
    cmd = [nvcc_path()] + ARCH_FLAGS + [
        '-lineinfo', '-O3', '-std=c++17', '-shared', '-Xcompiler', '-fPIC',
        '-cudart', 'static', '-o', tmp_lib, RUNTIME_SRC, '-ldl']
    logger.info('Building %s', RUNTIME_LIB)
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode == 0:
        # several processes (one per GPU) may build at once on a cold tree:
        # nobody ever sees a half-written library
        os.replace(tmp_lib, RUNTIME_LIB)
    elif os.path.exists(tmp_lib):
        os.remove(tmp_lib)
    if verbose:
        print(proc.stdout)
        print(proc.stderr)
    if proc.returncode != 0:
        raise ImportError('Unable to build the opty_b200 runtime library, '
                          'compilation failed. STDERR output from '
                          'compilation:\n{}'.format(proc.stderr))
    return RUNTIME_LIB


def _header_digest(source=''):
    hasher = hashlib.sha256()
    for name in ('colloc_kernel.cuh', 'colloc_params.h'):
        with open(os.path.join(CSRC, name), 'rb') as f:
            hasher.update(f.read())
    return hasher.hexdigest()


def module_flags(fmad=False, maxrregcount=None, opt_level=3):
    flags = ARCH_FLAGS + ['-cubin', '-O{}'.format(opt_level), '-lineinfo',
                          '-std=c++17',
                          '--fmad={}'.format('true' if fmad else 'false'),
                          '-I', CSRC]
    if maxrregcount:
        flags += ['-maxrregcount', str(int(maxrregcount))]
    return flags


def compile_module(source, flags, cache_dir=None, show_compile_output=False,
                   keep_source=True):
    """Compiles emitted CUDA-C ``source`` to a cubin, with a content-addressed
    cache (key = sha256 of source + flags + kernel header).

    Returns ``(cubin_bytes, cubin_path, cache_hit)``.
    """
    cache_dir = cache_dir or default_cache_dir()
    hasher = hashlib.sha256()
    hasher.update(source.encode())
    hasher.update(' '.join(flags).encode())
    hasher.update(_header_digest(source).encode())
    key = hasher.hexdigest()[:32]
    os.makedirs(cache_dir, exist_ok=True)
    cubin_path = os.path.join(cache_dir, 'colloc_{}.cubin'.format(key))
    if os.path.exists(cubin_path) and os.path.getsize(cubin_path) > 0:
        logger.info('Skipped compile, %s loaded.', cubin_path)
        with open(cubin_path, 'rb') as f:
            return f.read(), cubin_path, True

    src_path = os.path.join(cache_dir, 'colloc_{}.cu'.format(key))
    text = '// opty_code_hash={}\n'.format(key) + source
    try:
        with open(src_path) as f:
            same = f.read() == text
    except OSError:
        same = False
    if not same:
        # written under a private name first: another rank compiling the same
        # module must never read a truncated source
        tmp_src = '{}.{}.tmp'.format(src_path, os.getpid())
        with open(tmp_src, 'w') as f:
            f.write(text)
        os.replace(tmp_src, src_path)
    tmp_out = tempfile.NamedTemporaryFile(
        dir=cache_dir, suffix='.cubin.tmp', delete=False)
    tmp_out.close()
    cmd = [nvcc_path()] + list(flags) + ['-o', tmp_out.name, src_path]
    if show_compile_output:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    logger.info('Compiling the collocation module %s', src_path)
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0 and 'Segmentation fault' in proc.stderr:
        # ptxas 12.9 occasionally crashes on very long straight-line bodies
        # under tight register caps at -O3; its -O2 pipeline gets through
        logger.warning('ptxas crashed at -O3, retrying with -Xptxas -O2')
        proc = subprocess.run(cmd[:1] + ['-Xptxas', '-O2'] + cmd[1:],
                              capture_output=True, text=True)
    if show_compile_output:
        print(proc.stdout)
        print(proc.stderr)
    else:
        logger.debug(proc.stdout)
        logger.debug(proc.stderr)
    if proc.returncode != 0 or os.path.getsize(tmp_out.name) == 0:
        try:
            os.unlink(tmp_out.name)
        except OSError:
            pass
        raise ImportError(
            'Unable to compile the generated CUDA module {}, compilation '
            'failed. STDERR output from compilation:\n{}'.format(
                src_path, proc.stderr))
    os.replace(tmp_out.name, cubin_path)
    if not keep_source:
        os.unlink(src_path)
    with open(cubin_path, 'rb') as f:
        return f.read(), cubin_path, False


def expression_digest(hasher, expr):
    """Feeds a canonical, process-independent serialisation of a SymPy
    expression tree into ``hasher`` (pre-order, explicit stack: the discrete
    EOM of large models are millions of nodes deep in places).  Symbols and
    undefined functions contribute their names (and assumptions that change
    evaluation: none do), numbers their exact value."""
    import sympy as sm
    stack = [expr]
    up = hasher.update
    while stack:
        e = stack.pop()
        if isinstance(e, sm.Symbol):
            up(b'S' + e.name.encode() + b';')
        elif isinstance(e, sm.Integer):
            up(b'I' + str(int(e)).encode() + b';')
        elif isinstance(e, sm.Rational):
            up(b'Q' + str(e.p).encode() + b'/' + str(e.q).encode() + b';')
        elif isinstance(e, sm.Float):
            up(b'F' + repr(float(e)).encode() + b'@' +
               str(e._prec).encode() + b';')
        elif not e.args:
            up(b'A' + type(e).__name__.encode() + b':' + str(e).encode() +
               b';')
        else:
            name = getattr(getattr(e, 'func', None), '__name__',
                           type(e).__name__)
            up(b'(' + name.encode() + b':' + str(len(e.args)).encode() + b';')
            stack.extend(reversed(e.args))


def load_index(cache_dir, input_key):
    path = os.path.join(cache_dir or default_cache_dir(),
                        'index_{}.json'.format(input_key))
    if not os.path.exists(path):
        return None
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


def store_index(cache_dir, input_key, payload):
    cache_dir = cache_dir or default_cache_dir()
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, 'index_{}.json'.format(input_key))
    tmp = path + '.tmp{}'.format(os.getpid())
    with open(tmp, 'w') as f:
        json.dump(payload, f)
    os.replace(tmp, path)
