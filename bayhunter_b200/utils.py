"""utils with BayHunter's interface for what an inversion script touches (src/utils.py):
`load_params` (the `.ini` files of tutorial/config.ini / defaults/defaults.ini), `save_config`,
`save_baywatch_config`, `read_config`.  The INI dialect is configobj's as the reference uses it:
`[section]` headers, `key = value` lines, comma separated values are lists, every value except
`station` / `savepath` is evaluated as a Python expression (string_decode, src/utils.py:44-56)."""
import pickle
import os.path as op

KEYWORDS = ('station', 'savepath')


def _parse_ini(initfile):
    sections, cur = [], None
    with open(initfile) as fh:
        for raw in fh:
            line = raw.split('#', 1)[0].strip()
            if not line:
                continue
            if line.startswith('[') and line.endswith(']'):
                cur = {}
                sections.append((line[1:-1].strip(), cur))
                continue
            key, val = [s.strip() for s in line.split('=', 1)]
            cur[key] = val
    return sections


def _literal(text):
    """Value of an .ini entry: numbers, None / True / False, strings, tuples / lists and arithmetic on numbers
    ("(2048 * 16)", "1e-5", "2, 5").  The reference hands the text to eval() (src/utils.py:44-56); a parameter file
    needs none of what that allows beyond this."""
    import ast
    import operator as op_
    ops = {ast.Add: op_.add, ast.Sub: op_.sub, ast.Mult: op_.mul, ast.Div: op_.truediv, ast.Pow: op_.pow,
           ast.FloorDiv: op_.floordiv, ast.Mod: op_.mod}

    def ev(n):
        if isinstance(n, ast.Constant) and (n.value is None or isinstance(n.value, (int, float, str, bool))):
            return n.value
        if isinstance(n, (ast.Tuple, ast.List)):
            seq = [ev(e) for e in n.elts]
            return tuple(seq) if isinstance(n, ast.Tuple) else seq
        if isinstance(n, ast.UnaryOp) and isinstance(n.op, (ast.USub, ast.UAdd)):
            v = ev(n.operand)
            return -v if isinstance(n.op, ast.USub) else +v
        if isinstance(n, ast.BinOp) and type(n.op) in ops:
            a, b = ev(n.left), ev(n.right)
            if not all(isinstance(x, (int, float)) and not isinstance(x, bool) for x in (a, b)):
                raise ValueError("arithmetic on non-numbers in %r" % text)
            return ops[type(n.op)](a, b)
        raise ValueError("unsupported expression in parameter file: %r" % text)

    return ev(ast.parse(text.strip(), mode="eval").body)


def _decode(key, val):
    if key in KEYWORDS:
        if len(val) >= 2 and val[0] == val[-1] and val[0] in "'\"":
            return val[1:-1]
        return val
    try:                                    # "(2048 * 16)", "None", "1e-5", "0.9"
        out = _literal(val)
    except (SyntaxError, ValueError):
        out = [_literal(v) for v in val.split(',') if v.strip()]
    if isinstance(out, tuple) and ',' in val and not val.startswith('('):
        out = list(out)                     # configobj hands "a, b" over as a list of strings
    return out


def load_params(initfile):
    """[modelpriors, initparams] dictionaries of an ini file (src/utils.py:59-70)."""
    params = []
    for name, sec in _parse_ini(initfile):
        if name == 'datapaths':
            continue
        params.append({k: _decode(k, v) for k, v in sec.items()})
    return params


def save_baywatch_config(targets, path='.', priors=dict(), initparams=dict(), refmodel=dict()):
    """src/utils.py:102-124."""
    for target in targets.targets:
        target.get_covariance = None
    data = {'targets': targets.targets, 'priors': priors, 'initparams': initparams, 'refmodel': refmodel}
    with open(op.join(path, 'baywatch.pkl'), 'wb') as f:
        pickle.dump(data, f)


def save_config(targets, configfile, priors=dict(), initparams=dict()):
    """src/utils.py:127-153."""
    from .mcmcOptimizer import save_config as _save
    _save(targets, configfile, priors=priors, initparams=initparams)


def read_config(configfile):
    """src/utils.py:156-164."""
    with open(configfile, 'rb') as f:
        try:
            return pickle.load(f)
        except UnicodeDecodeError:
            f.seek(0)
            return pickle.load(f, encoding='latin1')
