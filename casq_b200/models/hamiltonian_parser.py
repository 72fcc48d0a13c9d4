"""Hamiltonian string-spec parser (the product's replacement for qiskit-dynamics'
``parse_backend_hamiltonian_dict``, called by casq at src/casq/models/hamiltonian_model.py:85-90).

A small tokenizer + recursive-descent evaluator over values that are either scalars or dense operators on
the extracted subsystems.  Grammar per term (after ``_SUM[i,lo,hi,...]`` expansion):
    term   := expr [ '||' CHANNEL ]
    expr   := prod (('+'|'-') prod)*
    prod   := unary (('*'|'/') unary)*
    unary  := '-' unary | atom
    atom   := NUMBER | VARIABLE | OPERATOR | '(' expr ')'
OPERATOR is one of I X Y Z Sp Sm O N A C followed by a subsystem label.  Subsystem 0 is the
least-significant Kronecker factor; a term touching a subsystem outside ``subsystem_list`` is dropped;
terms on the same channel are summed; channels come back lower-cased and sorted.
"""
from __future__ import annotations

import re
from typing import Optional

import numpy as np

from ..common.exceptions import CasqError

_OPS = ("Sp", "Sm", "I", "X", "Y", "Z", "O", "N", "A", "C")
_TOKEN = re.compile(r"\s*(?:(\d+\.?\d*(?:[eE][-+]?\d+)?)|([A-Za-z_][A-Za-z_0-9]*)|(.))")
_SUM = re.compile(r"^_SUM\[\s*(\w+)\s*,\s*(-?\d+)\s*,\s*(-?\d+)\s*,(.*)\]$")
_OPNAME = re.compile(r"^(Sp|Sm|I|X|Y|Z|O|N|A|C)(\d+)$")


def _ladder(dim: int) -> dict:
    a = np.zeros((dim, dim), dtype=complex)
    for n in range(1, dim):
        a[n - 1, n] = np.sqrt(n)
    ad = a.conj().T
    num = ad @ a
    eye = np.eye(dim, dtype=complex)
    return {"Sm": a, "A": a, "Sp": ad, "C": ad, "O": num, "N": num, "X": a + ad, "Y": -1j * (a - ad),
            "Z": eye - 2.0 * num, "I": eye}


def expand_sums(terms) -> list:
    out = []
    for term in terms:
        term = term.strip()
        m = _SUM.match(term)
        if not m:
            out.append(term)
            continue
        var, lo, hi, body = m.group(1), int(m.group(2)), int(m.group(3)), m.group(4)
        out.extend(expand_sums([body.replace("{" + var + "}", str(i)) for i in range(lo, hi + 1)]))
    return out


class _Dropped(Exception):
    """Raised while evaluating a term that touches a subsystem outside the extracted list."""


class _Evaluator:
    def __init__(self, text: str, variables: dict, dims: dict, order: list):
        self.toks = [(num, ident, sym) for num, ident, sym in _TOKEN.findall(text) if num or ident or sym.strip()]
        self.pos = 0
        self.vars = variables
        self.dims = dims
        self.order = order
        self.total = int(np.prod([dims[q] for q in order])) if order else 1

    def _peek(self):
        return self.toks[self.pos] if self.pos < len(self.toks) else None

    def _next(self):
        tok = self._peek()
        self.pos += 1
        return tok

    def _embed(self, name: str, q: int) -> np.ndarray:
        if q not in self.dims:
            raise _Dropped()
        local = _ladder(self.dims[q])[name]
        out = np.ones((1, 1), dtype=complex)
        for s in self.order:  # ascending labels, each new one to the LEFT
            out = np.kron(local if s == q else np.eye(self.dims[s], dtype=complex), out)
        return out

    @staticmethod
    def _mul(a, b):
        if isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
            return a @ b
        return a * b

    def expr(self):
        val = self.prod()
        while (tok := self._peek()) and tok[2] in "+-" and tok[2]:
            self._next()
            rhs = self.prod()
            val, rhs = self._lift(val, rhs)
            val = val + rhs if tok[2] == "+" else val - rhs
        return val

    def _lift(self, a, b):
        if isinstance(a, np.ndarray) and not isinstance(b, np.ndarray):
            b = b * np.eye(self.total, dtype=complex)
        elif isinstance(b, np.ndarray) and not isinstance(a, np.ndarray):
            a = a * np.eye(self.total, dtype=complex)
        return a, b

    def prod(self):
        val = self.unary()
        while (tok := self._peek()) and tok[2] in "*/" and tok[2]:
            self._next()
            rhs = self.unary()
            if tok[2] == "*":
                val = self._mul(val, rhs)
            else:
                if isinstance(rhs, np.ndarray):
                    raise CasqError("Hamiltonian string: division by an operator is not defined.")
                val = val / rhs
        return val

    def unary(self):
        tok = self._peek()
        if tok and tok[2] == "-":
            self._next()
            return -self.unary()
        if tok and tok[2] == "+":
            self._next()
            return self.unary()
        return self.atom()

    def atom(self):
        tok = self._next()
        if tok is None:
            raise CasqError("Hamiltonian string: unexpected end of term.")
        num, ident, sym = tok
        if num:
            return float(num)
        if ident:
            m = _OPNAME.match(ident)
            if m and ident not in self.vars:
                return self._embed(m.group(1), int(m.group(2)))
            if ident in self.vars:
                return float(self.vars[ident])
            if ident == "pi":
                return float(np.pi)
            raise CasqError(f"Hamiltonian string: unknown symbol '{ident}'.")
        if sym == "(":
            val = self.expr()
            closing = self._next()
            if closing is None or closing[2] != ")":
                raise CasqError("Hamiltonian string: missing ')'.")
            return val
        raise CasqError(f"Hamiltonian string: unexpected token '{sym}'.")


def parse_hamiltonian_dict(hamiltonian_dict: dict, subsystem_list: Optional[list] = None):
    """-> (static_operator [d,d], operators [K,d,d], channels list[str], subsystem_dims dict)."""
    for key in ("h_str", "qub", "vars"):
        if key not in hamiltonian_dict:
            raise CasqError(f"hamiltonian_dict is missing the '{key}' entry.")
    qub = {int(k): int(v) for k, v in hamiltonian_dict["qub"].items()}
    order = sorted(qub) if subsystem_list is None else sorted(int(q) for q in subsystem_list)
    for q in order:
        if q not in qub:
            raise CasqError(f"Extracted qubit {q} is not in the Hamiltonian ('qub' has {sorted(qub)}).")
    dims = {q: qub[q] for q in order}
    total = int(np.prod([dims[q] for q in order]))
    static = np.zeros((total, total), dtype=complex)
    by_channel: dict = {}
    for term in expand_sums(hamiltonian_dict["h_str"]):
        body, _, chan = term.partition("||")
        chan = chan.strip().lower() or None
        ev = _Evaluator(body, hamiltonian_dict["vars"], dims, order)
        try:
            val = ev.expr()
        except _Dropped:
            continue
        if ev.pos != len(ev.toks):
            raise CasqError(f"Hamiltonian string: trailing tokens in term '{term}'.")
        if not isinstance(val, np.ndarray):
            val = val * np.eye(total, dtype=complex)
        if chan is None:
            static = static + val
        else:
            by_channel[chan] = by_channel.get(chan, 0) + val
    channels = sorted(by_channel)
    operators = np.zeros((len(channels), total, total), dtype=complex)
    for k, c in enumerate(channels):
        operators[k] = by_channel[c]
    return static, operators, channels, dims
