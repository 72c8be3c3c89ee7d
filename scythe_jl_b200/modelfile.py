"""Reader for the reference's model-parameter files (``models/*.jl``), so that ``run_Scythe.jl``'s argument -- the path of
a Julia file whose body is ``model = ModelParameters(...)`` -- works unmodified with the B200 launcher (`run.py`).

Julia is not installed in this image, so the file is not executed: the restricted expression language the reference's
model files use (/root/reference/models/LinearAdvection1D.jl:1-22, /root/reference/models/cha_bell2024/*.jl) is parsed
directly -- constructor calls with keyword arguments, ``Dict(k => v, ...)``, strings, numbers, ``:symbols``,
``true``/``false`` and dotted constants such as ``CubicBSpline.R1T0`` / ``Chebyshev.R0``.  Anything else is rejected
with the line number (no silent guess).  A ``.py`` model file that defines ``model`` is accepted as well.
"""
from __future__ import annotations

import re
import runpy

from . import api

_TOKEN = re.compile(r"""
    (?P<ws>\s+|\#[^\n]*)
  | (?P<num>[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eEf][-+]?\d+)?)
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<sym>:[A-Za-z_α-ωΑ-Ω][A-Za-z_0-9α-ωΑ-Ω!]*)
  | (?P<name>[A-Za-z_α-ωΑ-Ω][A-Za-z_0-9α-ωΑ-Ω!]*(?:\.[A-Za-z_α-ωΑ-Ω][A-Za-z_0-9α-ωΑ-Ω!]*)*)
  | (?P<arrow>=>)
  | (?P<op>[(),=\[\]])
""", re.X)


class ModelFileError(ValueError):
    pass


def _tokens(text: str):
    pos, line = 0, 1
    out = []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise ModelFileError(f"line {line}: cannot read {text[pos:pos + 20]!r}")
        kind = m.lastgroup
        if kind != "ws":
            out.append((kind, m.group(), line))
        line += m.group().count("\n")
        pos = m.end()
    out.append(("eof", "", line))
    return out


_CONSTANTS = {"CubicBSpline": api.CubicBSpline, "Chebyshev": api.Chebyshev}


class _Parser:
    def __init__(self, text):
        self.t = _tokens(text)
        self.i = 0

    def peek(self):
        return self.t[self.i]

    def take(self, kind=None, value=None):
        k, v, ln = self.t[self.i]
        if (kind and k != kind) or (value is not None and v != value):
            raise ModelFileError(f"line {ln}: expected {value or kind}, found {v!r}")
        self.i += 1
        return v

    def value(self):
        k, v, ln = self.peek()
        if k == "num":
            self.take()
            txt = v.replace("f", "e")
            return float(txt) if any(c in txt for c in ".eE") else int(txt)
        if k == "str":
            self.take()
            return v[1:-1].replace('\\"', '"').replace("\\\\", "\\")
        if k == "sym":
            self.take()
            return v[1:]                       # physical_params / options are looked up by bare name on the device side
        if k == "op" and v == "[":
            self.take()
            items = []
            while not (self.peek()[0] == "op" and self.peek()[1] == "]"):
                items.append(self.value())
                if self.peek()[1] == ",":
                    self.take()
            self.take("op", "]")
            return items
        if k == "name":
            self.take()
            if self.peek()[0] == "op" and self.peek()[1] == "(":
                return self.call(v, ln)
            if v == "true":
                return True
            if v == "false":
                return False
            if v in ("pi", "π"):
                import math
                return math.pi
            head, _, attr = v.partition(".")
            if head in _CONSTANTS and attr and hasattr(_CONSTANTS[head], attr):
                c = getattr(_CONSTANTS[head], attr)
                return dict(c) if isinstance(c, dict) else c
            raise ModelFileError(f"line {ln}: unknown name {v!r}")
        raise ModelFileError(f"line {ln}: unexpected {v!r}")

    def call(self, name, ln):
        self.take("op", "(")
        pos, kw, pairs = [], {}, []
        while not (self.peek()[0] == "op" and self.peek()[1] == ")"):
            k, v, kl = self.peek()
            if k == "eof":
                raise ModelFileError(f"line {kl}: expected ) to close {name}( of line {ln}")
            nk, nv, _ = self.t[self.i + 1]
            if k == "name" and nk == "op" and nv == "=":
                self.take()
                self.take("op", "=")
                kw[v] = self.value()
            else:
                a = self.value()
                if self.peek()[0] == "arrow":
                    self.take()
                    pairs.append((a, self.value()))
                else:
                    pos.append(a)
            if self.peek()[1] == ",":
                self.take()
        self.take("op", ")")
        short = name.rsplit(".", 1)[-1]
        if short not in ("Dict", "GridParameters", "ModelParameters", "ChebyshevParameters"):
            raise ModelFileError(f"line {ln}: unsupported constructor {name}")
        if short == "Dict":
            if pos or kw:
                raise ModelFileError(f"line {ln}: Dict takes key => value pairs")
            return dict(pairs)
        if pairs or pos:
            raise ModelFileError(f"line {ln}: {name} takes keyword arguments")
        if short == "GridParameters":
            return api.GridParameters(**kw)
        if short == "ModelParameters":
            if "options" in kw:                # same defaults as src/Scythe.jl:17-20, user entries on top
                opts = {"semiimplicit": False, "exact_reference_state": False}
                opts.update(kw["options"])
                kw["options"] = opts
            return api.ModelParameters(**kw)
        if short == "ChebyshevParameters":
            return api.ChebyshevParameters(**kw)
        raise ModelFileError(f"line {ln}: unsupported constructor {name}")

    def assignments(self):
        env = {}
        while self.peek()[0] != "eof":
            name = self.take("name")
            self.take("op", "=")
            env[name] = self.value()
        return env


def parse_model_text(text: str) -> api.ModelParameters:
    env = _Parser(text).assignments()
    if "model" not in env or not isinstance(env["model"], api.ModelParameters):
        raise ModelFileError("the model file must assign `model = ModelParameters(...)`")
    return env["model"]


def load_model_file(path: str) -> api.ModelParameters:
    """``include(modelfile)`` of run_Scythe.jl:44 for the restricted syntax above (or a Python file defining ``model``)."""
    if str(path).endswith(".py"):
        env = runpy.run_path(str(path))
        if not isinstance(env.get("model"), api.ModelParameters):
            raise ModelFileError(f"{path}: no `model = ModelParameters(...)`")
        return env["model"]
    with open(path, encoding="utf-8") as f:
        return parse_model_text(f.read())
