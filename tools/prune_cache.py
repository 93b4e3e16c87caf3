"""Development aid: shrinks the in-tree compiled-module cache
(opty_b200/_cache) to what ``__graft_entry__.build()`` and the prepared
config-5 dumps reference.  Every option sweep leaves cubins behind; the cache
travels to the GPU box with the repository snapshot.

    python tools/prune_cache.py [--dry-run]
"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opty_b200 import build  # noqa: E402

CACHE = build.default_cache_dir()
keep = set()

real_load, real_compile = build.load_index, build.compile_module


def load_index(cache_dir, key):
    idx = real_load(cache_dir, key)
    if idx:
        keep.add('index_{}.json'.format(key))
        keep.add(idx['cubin'])
        keep.update(idx['extra'])
    return idx


def compile_module(*args, **kwargs):
    out = real_compile(*args, **kwargs)
    keep.add(os.path.basename(out[1]))
    return out


real_store = build.store_index


def store_index(cache_dir, key, payload):
    keep.add('index_{}.json'.format(key))
    keep.add(payload['cubin'])
    keep.update(payload['extra'])
    return real_store(cache_dir, key, payload)


build.load_index, build.compile_module = load_index, compile_module
build.store_index = store_index
import __graft_entry__  # noqa: E402
__graft_entry__.build()
for dump in glob.glob(os.path.join(CACHE, 'config5_*.json')):
    keep.add(os.path.basename(dump))
    with open(dump) as f:
        d = json.load(f)
    keep.add(os.path.basename(d['cubin']))
    keep.update(os.path.basename(e['cubin_path'])
                for e in d['meta'].get('extra_modules', ()))
keep.update(f for f in os.listdir(CACHE) if f.endswith('.pkl'))
removed = freed = 0
for name in os.listdir(CACHE):
    base = name[:-3] + '.cubin' if name.endswith('.cu') else name
    if name in keep or (name.endswith('.cu') and base in keep):
        continue
    path = os.path.join(CACHE, name)
    freed += os.path.getsize(path)
    removed += 1
    if '--dry-run' not in sys.argv:
        os.remove(path)
print('kept {} files, removed {} ({:.0f} MB)'.format(
    len(keep), removed, freed / 1e6))
