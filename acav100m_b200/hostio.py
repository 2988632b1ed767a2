"""Small host-side helpers shared by the two CLI shims: dotted-key config overrides, an attribute dict
that reads unknown keys as None (the reference wraps its config in ``DefaultMunch(None, ...)``,
clustering/code/args.py:73-83), fire-style ``--a.b.c=v`` parsing, brace expansion of shard globs,
pickle / json helpers and the run-id used in ``log_*.json`` (clustering/code/utils.py:55-69).
None of the reference's third-party CLI dependencies (fire, munch, braceexpand) is required."""
import ast
import copy
import datetime
import json
import os
import pickle
import platform
import re
import time
from pathlib import Path


class AttrDict(dict):
    """dict with attribute access; missing keys read as None (DefaultMunch(None) semantics)."""

    def __getattr__(self, key):
        if key.startswith('__'):
            raise AttributeError(key)
        return self.get(key)

    def __setattr__(self, key, value):
        self[key] = value

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def objectify(tree):
    if isinstance(tree, dict):
        return AttrDict({k: objectify(v) for k, v in tree.items()})
    return tree


def file_stem(filename):
    """``Path(filename).stem`` for the per-clip loops (pathlib builds a path object per call: ~10 us each, minutes at
    10^8 clips): last path component without its final suffix; a leading dot or a trailing dot is not a suffix."""
    name = str(filename)
    if not name or name[-1] in '/\\' or '\\' in name:
        return Path(name).stem
    name = name.rsplit('/', 1)[-1]
    if name in ('.', '..'):
        return Path(str(filename)).stem
    i = name.rfind('.')
    return name[:i] if 0 < i < len(name) - 1 else name


def update_args(args, overrides):
    """Merge ``{'a.b.c': v}`` into a nested dict (reference args.py `_update_args`)."""
    for key, value in overrides.items():
        node = args
        parts = key.split('.')
        for part in parts[:-1]:
            if not isinstance(node.get(part), dict):
                node[part] = {}
            node = node[part]
        node[parts[-1]] = value
    return args


def _literal(text):
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError):
        return text


def parse_cli(argv):
    """``cmd --k=v --a.b=c --flag`` -> (cmd, {k: v, 'a.b': c, flag: True}) like python-fire."""
    command, kwargs = None, {}
    i, argv = 0, list(argv)
    while i < len(argv):
        tok = argv[i]
        i += 1
        if tok.startswith('--'):
            body = tok[2:]
            if '=' in body:
                key, val = body.split('=', 1)
                kwargs[key] = _literal(val)
            elif i < len(argv) and not argv[i].startswith('--'):
                kwargs[body] = _literal(argv[i])
                i += 1
            else:
                kwargs[body] = True
        elif command is None:
            command = tok
        else:
            raise SystemExit("unexpected argument: %s" % tok)
    return command, kwargs


_BRACE = re.compile(r'\{([^{}]*)\}')


def braceexpand(pattern):
    """``shard-{000000..000003}.pkl`` / ``{a,b}`` expansion (bash semantics for the forms the
    reference's shard paths use)."""
    pattern = str(pattern)
    m = _BRACE.search(pattern)
    if not m:
        return [pattern]
    body = m.group(1)
    rng = re.fullmatch(r'(-?\d+)\.\.(-?\d+)', body)
    if rng:
        lo, hi = rng.group(1), rng.group(2)
        width = max(len(lo), len(hi)) if (lo.startswith('0') or hi.startswith('0')) and len(lo) > 1 else 0
        a, b = int(lo), int(hi)
        step = 1 if b >= a else -1
        options = [str(v).zfill(width) for v in range(a, b + step, step)]
    else:
        options = body.split(',')
    out = []
    for opt in options:
        out.extend(braceexpand(pattern[:m.start()] + opt + pattern[m.end():]))
    return out


def load_pickle(path):
    with open(path, 'rb') as f:
        return pickle.load(f)


def dump_pickle(data, path):
    with open(path, 'wb') as f:
        pickle.dump(data, f)


def load_json(path):
    with open(path) as f:
        return json.load(f)


def dump_json(data, path, indent=None):
    with open(path, 'w') as f:
        json.dump(data, f, indent=indent)


def get_run_info():
    return {'hostname': platform.uname()[1], 'pid': os.getpid(), 'timestamp': int(time.time()),
            'time': str(datetime.datetime.now())}


def get_run_id(run_info=None):
    run_info = run_info or get_run_info()
    return '_'.join(str(run_info[k]) for k in ('hostname', 'pid', 'timestamp') if k in run_info)


def resolve_paths(tree, root):
    """reference args.py `process_paths`: every 'path' is made absolute, '*_file'/'*_dir' join root."""
    if 'path' in tree and tree['path'] is not None:
        tree['path'] = Path(tree['path']).resolve()
    for key, val in list(tree.items()):
        if isinstance(val, dict):
            resolve_paths(val, root)
        elif val is not None and (key.endswith('_file') or key.endswith('_dir')):
            tree[key] = root / val
    return tree
