"""Collect the numeric constants of Bogaert's asymptotic Gauss-Legendre
formulas (n > 100) into jaxfun_b200/data/fastgl_coeffs.json.

They are published fit constants (I. Bogaert, "Iteration-free computation of
Gauss-Legendre quadrature nodes and weights", SIAM J. Sci. Comput. 36(3), 2014)
and cannot be re-derived, so this script reads them as *data* from the
reference checkout (`src/jaxfun/utils/fastgl.py:10-11, 232-508`) and stores
them as plain coefficient vectors, highest power first (Horner order).
Needs /root/reference; the JSON it writes is committed.
"""
import json
import os
import re

REF = "/root/reference/src/jaxfun/utils/fastgl.py"
OUT = os.path.join(os.path.dirname(__file__), "..", "jaxfun_b200", "data", "fastgl_coeffs.json")
NUM = r"[-+]?\s*\d+\.\d+(?:e[-+]?\d+)?"


def literals(block: str):
    return ["".join(s.split()) for s in re.findall(NUM, block)]


def main():
    src = open(REF).read()
    out = {}
    for name in ("J1", "JZ"):
        m = re.search(rf"^{name} = jnp\.array\(\[([^\]]*)\]\)", src, re.M)
        out[name] = [s.strip() for s in m.group(1).split(",")]
    # besseljzero / besselj1squared bodies
    kb20 = src[src.index("def kb20"): src.index("return jax.lax.cond(k > 20")]
    lits = literals(kb20)
    assert lits[0] == "-0.25" and lits[1] == "1.0", lits[:3]
    out["kb20"] = lits[2:]          # ascending powers of r^2: c0 (0.125) ... c8
    km21 = src[src.index("def km21"): src.index("return jax.lax.cond(k > 21")]
    lits = literals(km21)
    assert lits[0] == "1.0" and lits[1] == "-0.25", lits[:3]
    out["km21"] = lits[2:]          # c0, then coefficients of x2^2 * (c1 + x2*(c2 + ...))
    body = src[src.index("def GLPairS"): src.index("# Then refine with the paper expansions")]
    for name in ("SF1T", "SF2T", "SF3T", "WSF1T", "WSF2T", "WSF3T"):
        start = body.index(f"{name}: Array")
        nxt = min([body.index(f"{o}: Array") for o in ("SF1T", "SF2T", "SF3T", "WSF1T", "WSF2T", "WSF3T")
                   if body.find(f"{o}: Array") > start] + [len(body)])
        out[name] = literals(body[start:nxt])  # Horner order, highest power first
    # store as decimal strings (exact published digits); consumers call float()
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, len(v), v[0], v[-1])


if __name__ == "__main__":
    main()
