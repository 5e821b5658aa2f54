"""Run an UNMODIFIED BER driver script of the reference on the B200 engine.

    python -m informationbottleneckdecodingldpc_b200.run_driver Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py
    python -m informationbottleneckdecodingldpc_b200.run_driver <driver.py> --set min_errors=500 --set EbN0_dB_max_value=1.0

The drivers import their classes by module path (``from Discrete_LDPC_decoding.discrete_LDPC_decoder import
Discrete_LDPC_Decoder_class``, ``from AWGN_Channel_Transmission.AWGN_Quantizer_BPSK import ...``,
Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py:8-10).  ``install_shadow()`` registers this package's
sub-packages under exactly those top-level names in ``sys.modules``, so the driver source runs as it is: quantizer,
transmitter, channel, encoder and all four decoder classes resolve to the B200 back-end.  ``install_stubs()`` provides
no-op ``matplotlib`` / ``matplotlib.pyplot`` modules where matplotlib is not installed (the drivers only plot the
final curve).  The script is executed with its own directory as working directory (its relative paths to
``LDPC_codes/`` and the decoder-config pickle stay valid).

``--set name=value`` overrides a plain ``name = <literal>`` assignment of the script (e.g. ``min_errors``,
``EbN0_dB_max_value``) to shorten a Monte-Carlo run; without it the source is executed byte for byte.
"""
from __future__ import annotations

import importlib
import os
import re
import sys
import types

_PKG = __name__.rsplit(".", 1)[0]

# reference top-level package -> modules the drivers and each other import from it
SHADOWED = {
    "AWGN_Channel_Transmission": ["AWGN_Quantizer_BPSK", "AWGN_channel", "LDPC_Transmitter"],
    "Discrete_LDPC_decoding": ["discrete_LDPC_decoder", "discrete_LDPC_decoder_irreg", "LDPC_encoder"],
    "Continous_LDPC_Decoding": ["min_sum_decoder_irreg", "bp_decoder_irreg"],
}


def install_shadow() -> None:
    """Alias the reference's module paths to this package (idempotent)."""
    for top, mods in SHADOWED.items():
        pkg = importlib.import_module(f"{_PKG}.{top}")
        sys.modules[top] = pkg
        for m in mods:
            sys.modules[f"{top}.{m}"] = importlib.import_module(f"{_PKG}.{top}.{m}")


def install_stubs() -> None:
    """No-op matplotlib when the real one is missing; the drivers call mpl.use / rcParams.update / a few pyplot
    functions after the simulation."""
    try:
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
        matplotlib.use("Agg", force=True)
        _use = matplotlib.use
        matplotlib.use = lambda *a, **k: None      # the drivers ask for the "pgf" backend (needs a LaTeX install)
        matplotlib.rcParams.update = lambda *a, **k: None
        del _use
        return
    except Exception:
        pass

    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return lambda *a, **k: None

    mpl = _Anything("matplotlib")
    mpl.rcParams = {}
    plt = _Anything("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


class Silent:
    """Object whose every attribute is a no-op callable."""

    def __getattr__(self, name):
        return lambda *a, **k: None


def apply_overrides(source: str, overrides: dict) -> str:
    """Replace the right-hand side of the first ``name = <literal>`` line for every override."""
    for name, value in overrides.items():
        pat = re.compile(rf"^(?P<ind>[ \t]*){re.escape(name)}[ \t]*=[ \t]*[^=\n][^\n]*$", re.M)
        source, n = pat.subn(lambda m: f"{m.group('ind')}{name} = {value!r}", source, count=1)
        if n != 1:
            raise KeyError(f"driver has no plain assignment to {name!r}")
    return source


def run(path: str, overrides: dict | None = None, chdir: bool = True, inject: dict | None = None) -> dict:
    """Execute the driver at ``path``; returns its global namespace (BER_vector, EbN0_dB_vector, ...).
    ``inject`` pre-defines globals the script uses without defining them (two of the shipped drivers end with
    ``pb.push_note(...)`` on an undefined Pushbullet client; ``inject={"pb": run_driver.Silent()}`` lets them finish)."""
    install_shadow()
    install_stubs()
    path = os.path.abspath(path)
    with open(path) as fh:
        source = fh.read()
    if overrides:
        source = apply_overrides(source, overrides)
    ns = {"__name__": "__main__", "__file__": path, "__builtins__": __builtins__}
    ns.update(inject or {})
    old = os.getcwd()
    import numpy as np
    err = np.geterr()
    try:
        if chdir:
            os.chdir(os.path.dirname(path))
        exec(compile(source, path, "exec"), ns)
    finally:
        os.chdir(old)
        np.seterr(**err)          # the drivers set np.seterr(all='raise')
    return ns


def main(argv=None) -> int:
    import argparse
    import ast
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("driver")
    ap.add_argument("--set", action="append", default=[], metavar="name=value")
    a = ap.parse_args(argv)
    ov = {}
    for item in a.set:
        k, _, v = item.partition("=")
        try:
            ov[k.strip()] = ast.literal_eval(v)
        except Exception:
            ov[k.strip()] = v
    run(a.driver, ov, inject={"pb": Silent()})
    return 0


if __name__ == "__main__":
    sys.exit(main())
