"""Operation count of one env step of the CPU oracle (TEST INFRASTRUCTURE ONLY, see oracle/threefry.py header).

SURVEY.md §8(d) fixes the algorithmic figure bench.py reports against: 4.30e4 FLOPs per env-step, counted on the
pipeline specialised to the Ant (identity inertia, +z ground normal, frozen bodies and exactly-culled wall pairs
removed; every add / sub / mul / div / sqrt / compare / select / atan2 = 1, an FMA = 2). This script re-derives a
count the same way -- by running the oracle's own `System.step` (oracle/brax_v1.py, the restatement of
brax.System.step as called at /root/reference/po_brax/envs/ant_heavenhell.py:108) on arrays that count their
element-wise operations -- for the oracle AS WRITTEN: all `nb` bodies incl. the frozen ones, general normals, masks
instead of branches. The as-written figure is therefore an upper bound of the specialised one; the check that
matters (tests/test_masks_and_oracle_envs.py) is that the bench's 4.30e4 does not exceed it, i.e. that
`roofline.achieved` cannot be inflated by the FLOP figure.

    python -m oracle.count_ops            # prints the per-kind counts for plain Ant (no walls) and HeavenHell
"""
from collections import Counter

import numpy as np

_ARITH = {'add': 'add', 'subtract': 'add', 'multiply': 'mul', 'negative': 'neg', 'divide': 'div', 'true_divide': 'div',
          'sqrt': 'sqrt', 'arctan2': 'atan2', 'minimum': 'cmp', 'maximum': 'cmp', 'less': 'cmp', 'less_equal': 'cmp',
          'greater': 'cmp', 'greater_equal': 'cmp', 'equal': 'cmp', 'not_equal': 'cmp', 'absolute': 'abs',
          'square': 'mul', 'sin': 'sincos', 'cos': 'sincos', 'sign': 'cmp', 'reciprocal': 'div', 'clip': 'cmp'}
_FREE = {'logical_and', 'logical_or', 'logical_not', 'logical_xor', 'isnan', 'isfinite', 'bitwise_and', 'bitwise_or',
         'invert', 'positive', 'copysign', 'floor', 'rint', 'trunc'}   # boolean / bit plumbing: not FP work
COUNTS = Counter()


def _plain(x):
    if isinstance(x, Counting):
        return x.view(np.ndarray)
    if isinstance(x, (list, tuple)):
        return type(x)(_plain(v) for v in x)
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    return x


def _wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, Counting):
        return x.view(Counting)
    if isinstance(x, (list, tuple)):
        return type(x)(_wrap(v) for v in x)
    return x


class Counting(np.ndarray):
    """ndarray that adds the number of produced elements to COUNTS[kind] for every arithmetic ufunc it takes part in."""

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        if 'out' in kw:
            kw['out'] = _plain(kw['out'])
        out = getattr(ufunc, method)(*_plain(inputs), **kw)
        name = ufunc.__name__
        if name in _ARITH:
            if method == 'reduce':     # a sum over k elements = k - 1 adds per output element
                n_in = int(np.size(_plain(inputs[0])))
                COUNTS[_ARITH[name]] += n_in - int(np.size(out))
            else:
                COUNTS[_ARITH[name]] += int(np.size(out))
        elif name not in _FREE:
            COUNTS['other:' + name] += int(np.size(out))
        return _wrap(out)

    def __array_function__(self, func, types, args, kwargs):
        out = func(*_plain(args), **_plain(kwargs))
        if func is np.where and len(args) == 3:
            COUNTS['select'] += int(np.size(out))
        elif func in (np.sum,):
            pass  # routed through add.reduce above
        elif func in (np.clip,):
            COUNTS['cmp'] += 2 * int(np.size(out))
        return _wrap(out)


def count_step(env_name='ant', walls=False, n=4, seed=0):
    """Per-env operation counts (dict kind -> count) of one `System.step` (10 substeps) of the oracle."""
    from . import envs as oenvs
    from . import threefry as tf
    kw = {} if env_name == 'ant' else {'walls': walls}
    env = oenvs.ENVS[env_name](**kw)
    keys = tf.split(tf.prng_key(seed), n)
    st = env.reset(keys)
    qp = st.qp
    qp = type(qp)(*(np.asarray(getattr(qp, f)).view(Counting) for f in ('pos', 'rot', 'vel', 'ang')))
    act = np.linspace(-1, 1, n * 8, dtype=np.float32).reshape(n, 8).view(Counting)
    COUNTS.clear()
    env.sys.step(qp, act)
    return {k: v / n for k, v in sorted(COUNTS.items())}


def flops(counts):
    """SURVEY §8(d) convention: every counted kind is 1 FLOP (an FMA would be a mul + an add = 2)."""
    return sum(v for k, v in counts.items() if not k.startswith('other:'))


if __name__ == '__main__':
    for name, walls in (('ant', False), ('ant_heavenhell', False), ('ant_heavenhell', True)):
        c = count_step(name, walls)
        print(f'{name} walls={walls}: {flops(c):.0f} ops per env-step as written ->', {k: round(v) for k, v in c.items()})
