"""Derived-expression parameters: ``Parameter(derived='{b}**2')`` (reference desilike/parameter.py:795-808 ``Parameter.eval``; placeholders
``{name}`` are other parameters, the expression is Python / NumPy arithmetic evaluated per point at reference base.py:537).

On the CUDA path the expression is compiled ONCE to a reverse-Polish program (opcodes of include/desilike_b200.h) that a small kernel runs per
point in front of every pass; nothing is ``eval``-ed.  Supported: + - * / ** unary minus, numeric constants, ``e`` / ``pi``, and the
functions exp, log, sqrt, log10, abs, sin, cos, tanh (bare, or as attributes of ``np`` / ``jnp`` / ``numpy`` / ``math``).  Anything else raises.
"""
import ast
import math
import re

import numpy as np

OPS = {'push': 0, 'const': 1, 'add': 2, 'sub': 3, 'mul': 4, 'div': 5, 'pow': 6, 'neg': 7, 'exp': 8, 'log': 9, 'sqrt': 10, 'log10': 11, 'abs': 12,
       'sin': 13, 'cos': 14, 'tanh': 15}
_BIN = {ast.Add: 'add', ast.Sub: 'sub', ast.Mult: 'mul', ast.Div: 'div', ast.Pow: 'pow'}
_FUNCS = {'exp', 'log', 'sqrt', 'log10', 'abs', 'sin', 'cos', 'tanh', 'absolute', 'fabs'}
_CONSTS = {'pi': math.pi, 'e': math.e}


class ExpressionError(ValueError):
    pass


def depends(expr):
    """Parameter names an expression refers to, in order of first appearance (reference parameter.py ``Parameter.depends``)."""
    out = []
    for name in re.findall(r'\{([^{}]+)\}', str(expr)):
        if name not in out: out.append(name)
    return out


def compile_expression(expr, columns):
    """``expr`` with ``{name}`` placeholders -> (program [(opcode, argument)], constants list); ``columns``: name -> column of the extended row."""
    names = depends(expr)
    src = str(expr)
    for i, name in enumerate(names):
        if name not in columns: raise ExpressionError('expression {!r} depends on unknown parameter {!r}'.format(expr, name))
        src = src.replace('{' + name + '}', '__p{:d}__'.format(i))
    try:
        tree = ast.parse(src, mode='eval').body
    except SyntaxError as exc:
        raise ExpressionError('cannot parse derived expression {!r}: {}'.format(expr, exc))
    prog, consts = [], []

    def const(value):
        value = float(value)
        if value not in consts: consts.append(value)
        prog.append((OPS['const'], consts.index(value)))

    def visit(node):
        if isinstance(node, ast.BinOp) and type(node.op) in _BIN:
            visit(node.left); visit(node.right); prog.append((OPS[_BIN[type(node.op)]], 0)); return
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            visit(node.operand)
            if isinstance(node.op, ast.USub): prog.append((OPS['neg'], 0))
            return
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
            const(node.value); return
        if isinstance(node, ast.Name):
            m = re.fullmatch(r'__p(\d+)__', node.id)
            if m: prog.append((OPS['push'], int(columns[names[int(m.group(1))]]))); return
            if node.id in _CONSTS: const(_CONSTS[node.id]); return
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in ('np', 'jnp', 'numpy', 'math') and node.attr in _CONSTS:
            const(_CONSTS[node.attr]); return
        if isinstance(node, ast.Call) and len(node.args) == 1 and not node.keywords:
            fn = node.func
            name = fn.id if isinstance(fn, ast.Name) else (fn.attr if isinstance(fn, ast.Attribute) and isinstance(fn.value, ast.Name)
                                                           and fn.value.id in ('np', 'jnp', 'numpy', 'math') else None)
            if name in _FUNCS:
                visit(node.args[0]); prog.append((OPS['abs' if name in ('absolute', 'fabs') else name], 0)); return
        raise ExpressionError('derived expression {!r}: construct {!r} is not on the CUDA path'.format(expr, ast.dump(node)[:80]))

    visit(tree)
    return prog, consts


def evaluate_host(expr, values):
    """Host evaluation of the same restricted grammar (plan-time use: default values of derived parameters)."""
    names = list(values)
    prog, consts = compile_expression(expr, {n: i for i, n in enumerate(names)})
    st = []
    for op, arg in prog:
        if op == 0: st.append(np.asarray(values[names[arg]], dtype='f8'))
        elif op == 1: st.append(np.float64(consts[arg]))
        elif op <= 6:
            b, a = st.pop(), st.pop()
            st.append([None, None, a + b, a - b, a * b, a / b, a**b][op] if op != 6 else np.power(a, b))
        else:
            a = st.pop()
            st.append({7: np.negative, 8: np.exp, 9: np.log, 10: np.sqrt, 11: np.log10, 12: np.abs, 13: np.sin, 14: np.cos, 15: np.tanh}[op](a))
    return st[0]
