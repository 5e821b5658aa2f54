#!/usr/bin/env python3
"""Build oracle/_ref/libref_kernels.so: the REFERENCE's own OpenCL kernel sources compiled
for the host CPU.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference cannot run as shipped
(no pyopencl / pocl / mako on the image), but its four ``.cl`` files are plain
OpenCL C 1.x and compile as C++17 once
  (i)  the Mako placeholders ``${ cn_degree }``, ``${ vn_degree }``, ``${ match }``,
       ``${ msg_at_time }``, ``${Nvar}`` are substituted, and
  (ii) a small prelude supplies ``kernel/global/local/private``, ``get_global_id`` ...
The sources are read where they lie under the reference tree, rendered into
``oracle/_ref/`` (git-ignored, never committed) and wrapped by a driver, written here, that
replays the reference's host loops (discrete_LDPC_decoder_irreg.py:245-341,
min_sum_decoder_irreg.py:221-287, bp_decoder_irreg.py:221-286, AWGN_Quantizer_BPSK.py:183-248).

``msg_at_time`` and ``Nvar`` are rendered to run-time globals so one library serves every
batch size; ``cn_degree``/``vn_degree``/``match`` stay compile-time (they size private arrays),
one namespace per requested configuration.

Work-group emulation: the NDRange is executed with the outer (node) dimension split across
OpenMP threads.  The regular ``varnode_update`` stages its LUT slice in ``local`` memory
behind a barrier (kernels_template.cl:122-130); each thread plays one work-group: its first
work-item of a launch performs the cooperative fill (lid=0, local_size=1), the remaining
ones see lid >= size and reuse the filled ``local`` array, exactly what a CPU OpenCL runtime
with one large work-group per core does.
"""
from __future__ import annotations

import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
DEFAULT_REF = "/root/reference"

# (family, cn_degree, vn_degree, match)
DEFAULT_CONFIGS = [
    ("reg", 6, 3, None),       # C1 (3,6)
    ("reg", 4, 2, None), ("reg", 4, 3, None),      # toy regular codes for golden vectors
    ("irr", 6, 3, False), ("irr", 5, 3, False),
    ("irr", 8, 11, True), ("irr", 8, 11, False),   # WLAN
    ("irr", 7, 8, True), ("irr", 7, 8, False),     # DVB-S2
    ("irr", 5, 4, True), ("irr", 5, 4, False),     # synthetic corner-case code (deg-1 VN, deg-2 CN)
    ("llr", 6, 3, None), ("llr", 8, 11, None), ("llr", 7, 8, None), ("llr", 5, 4, None),
]

PRELUDE = r"""
// ---- prelude written for this repository (not reference code) ----
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
static int g_msg_at_time = 1;
static int g_Nvar = 1;
static thread_local int t_gid[3] = {0, 0, 0};
static thread_local int t_lid0 = 0, t_lsz0 = 1;
static inline int get_global_id(int d) { return t_gid[d]; }
static inline int get_local_id(int d) { return d == 0 ? t_lid0 : 0; }
static inline int get_local_size(int d) { return d == 0 ? t_lsz0 : 1; }
#define CLK_LOCAL_MEM_FENCE 0
static inline void barrier(int) {}
static inline double sign(double x) { return (x > 0.0) - (x < 0.0); }
using std::min; using std::log; using std::exp;
#define kernel
#define __kernel
#define global
#define private
#define local static thread_local
// NDRange (n0 x n1), outer dimension over OpenMP threads, one emulated work-group per thread.
template <class F> static inline void ndrange(int n0, int n1, F f) {
#pragma omp parallel
  {
    bool first = true;
#pragma omp for schedule(static)
    for (int i = 0; i < n0; i++) {
      t_gid[0] = i;
      for (int j = 0; j < n1; j++) {
        t_gid[1] = j;
        if (first) { t_lid0 = 0; t_lsz0 = 1; first = false; }
        else       { t_lid0 = 1 << 30; t_lsz0 = 1 << 30; }
        f();
      }
    }
  }
}
struct ref_graph { int N, M, E; const int *sc, *dc, *tc, *sv, *dv, *tv; };
"""

HOSTLOOP_IB_REG = r"""
// host loop of discrete_LDPC_decoder.py:202-295 (regular kernels: no matching arguments)
static int decode(const ref_graph& g, int Tc, int T, int imax, const int* C, const int* V,
                  const int*, const int*, const int* ch, int B, int early, int* out) {
  g_msg_at_time = B;
  std::vector<int> cin((size_t)g.E * B), vin((size_t)g.E * B), syn((size_t)g.M * B);
  ndrange(g.N, B, [&] { send_channel_values_to_checknode_inbox(ch, g.sv, g.dv, g.tv, cin.data()); });
  ndrange(g.M, B, [&] { checknode_update_iter0(cin.data(), g.sc, g.dc, g.tc, vin.data(), Tc, T, C); });
  int i_num = 1; bool zero = false;
  while (i_num < imax && !zero) {
    ndrange(g.N, B, [&] { varnode_update(ch, vin.data(), g.sv, g.dv, g.tv, cin.data(), Tc, T, i_num - 1, V); });
    ndrange(g.M, B, [&] { checknode_update(cin.data(), g.sc, g.dc, g.tc, vin.data(), Tc, T, i_num - 1, C); });
    if (early) {
      ndrange(g.M, B, [&] { calc_syndrome(cin.data(), g.sc, g.dc, T, syn.data()); });
      long s = 0; for (int x : syn) s += x;
      if (s == 0) zero = true;
    }
    i_num++;
  }
  ndrange(g.N, B, [&] { calc_varnode_output(ch, vin.data(), g.sv, g.dv, Tc, T, i_num - 1, V, out); });
  return i_num;
}
"""

HOSTLOOP_IB_IRR = r"""
// host loop of discrete_LDPC_decoder_irreg.py:245-341
static int decode(const ref_graph& g, int Tc, int T, int imax, const int* C, const int* V,
                  const int* MC, const int* MV, const int* ch, int B, int early, int* out) {
  g_msg_at_time = B;
  std::vector<int> cin((size_t)g.E * B), vin((size_t)g.E * B), syn((size_t)g.M * B);
  ndrange(g.N, B, [&] { send_channel_values_to_checknode_inbox(ch, g.sv, g.dv, g.tv, cin.data()); });
  ndrange(g.M, B, [&] { checknode_update_iter0(cin.data(), g.sc, g.dc, g.tc, vin.data(), Tc, T, C, MC); });
  int i_num = 1; bool zero = false;
  while (i_num < imax && !zero) {
    ndrange(g.N, B, [&] { varnode_update(ch, vin.data(), g.sv, g.dv, g.tv, cin.data(), Tc, T, i_num - 1, V, MV); });
    ndrange(g.M, B, [&] { checknode_update(cin.data(), g.sc, g.dc, g.tc, vin.data(), Tc, T, i_num - 1, C, MC); });
    if (early) {
      ndrange(g.M, B, [&] { calc_syndrome(cin.data(), g.sc, g.dc, T, syn.data()); });
      long s = 0; for (int x : syn) s += x;
      if (s == 0) zero = true;
    }
    i_num++;
  }
  ndrange(g.N, B, [&] { calc_varnode_output(ch, vin.data(), g.sv, g.dv, Tc, T, i_num - 1, V, out); });
  return i_num;
}
"""

HOSTLOOP_LLR = r"""
// host loops of min_sum_decoder_irreg.py:221-287 (algo 0) and bp_decoder_irreg.py:221-286 (algo 1)
static int decode(const ref_graph& g, int algo, int imax, const double* ch, int B, int early, double* out) {
  g_msg_at_time = B;
  std::vector<double> cin((size_t)g.E * B), vin((size_t)g.E * B);
  std::vector<int> syn((size_t)g.M * B);
  ndrange(g.N, B, [&] { send_channel_values_to_checknode_inbox(ch, g.sv, g.dv, g.tv, cin.data()); });
  int i_num = 1; bool zero = false;
  while (i_num < imax && !zero) {
    if (algo == 0) ndrange(g.M, B, [&] { checknode_update_minsum(cin.data(), g.sc, g.dc, g.tc, vin.data()); });
    else           ndrange(g.M, B, [&] { checknode_update(cin.data(), g.sc, g.dc, g.tc, vin.data()); });
    ndrange(g.N, B, [&] { varnode_update(ch, vin.data(), g.sv, g.dv, g.tv, cin.data()); });
    if (early) {
      ndrange(g.M, B, [&] { calc_syndrome(cin.data(), g.sc, g.dc, syn.data()); });
      long s = 0; for (int x : syn) s += x;
      if (s == 0) zero = true;
    }
    i_num++;
  }
  ndrange(g.N, B, [&] { calc_varnode_output(ch, vin.data(), g.sv, g.dv, out); });
  return i_num;
}
"""

SOURCES = {
    "reg": "Discrete_LDPC_decoding/kernels_template.cl",
    "irr": "Discrete_LDPC_decoding/kernels_template_irreg.cl",
    "llr": "Continous_LDPC_Decoding/kernels_min_and_BP.cl",
    "qnt": "AWGN_Channel_Transmission/kernels_quanti_template.cl",
}


def render(text: str, **subst) -> str:
    def repl(m):
        key = m.group(1)
        if key not in subst:
            raise KeyError(f"unexpected placeholder {key}")
        return str(subst[key])
    return re.sub(r"\$\{\s*(\w+)\s*\}", repl, text)


def cfg_name(fam, dc, dv, match):
    return f"{fam}_{dc}_{dv}" + ("" if match is None else ("_m1" if match else "_m0"))


def build(ref_root: str = DEFAULT_REF, configs=None, quiet: bool = False) -> str:
    configs = list(configs or DEFAULT_CONFIGS)
    os.makedirs(OUT, exist_ok=True)
    texts = {k: open(os.path.join(ref_root, p)).read() for k, p in SOURCES.items()}
    parts = [PRELUDE]
    table_ib, table_llr = [], []
    for fam, dc, dv, match in configs:
        name = cfg_name(fam, dc, dv, match)
        inc = os.path.join(OUT, name + ".inc")
        subst = dict(cn_degree=dc, vn_degree=dv, msg_at_time="(g_msg_at_time)",
                     match="true" if match else "false")
        with open(inc, "w") as fh:
            fh.write(render(texts[fam], **subst))
        parts.append(f"#undef CN_DEGREE\n#undef VN_DEGREE\n#undef MATCH\n#undef LLR_MAX\n"
                     f"namespace {name} {{\n#include \"{name}.inc\"\n")
        parts.append({"reg": HOSTLOOP_IB_REG, "irr": HOSTLOOP_IB_IRR, "llr": HOSTLOOP_LLR}[fam])
        parts.append("}\n")
        if fam == "llr":
            table_llr.append((name, dc, dv))
        else:
            table_ib.append((name, 0 if fam == "reg" else 1, dc, dv, -1 if match is None else int(match)))
    qinc = os.path.join(OUT, "qnt.inc")
    with open(qinc, "w") as fh:
        fh.write(render(texts["qnt"], Nvar="(g_Nvar)"))
    parts.append('namespace qnt {\n#include "qnt.inc"\n}\n')
    # ---- C ABI of the shim (written here) ----
    api = ['extern "C" {']
    api.append("int ref_ib_decode(int irregular, int DC, int DV, int match, int N, int M, int E,"
               " const int* sc, const int* dc, const int* tc, const int* sv, const int* dv, const int* tv,"
               " int Tc, int T, int imax, const int* C, const int* V, const int* MC, const int* MV,"
               " const int* ch, int B, int early, int* out) {\n"
               "  ref_graph g{N, M, E, sc, dc, tc, sv, dv, tv};")
    for name, irr, dc, dv, match in table_ib:
        cond = f"irregular == {irr} && DC == {dc} && DV == {dv}" + ("" if match < 0 else f" && match == {match}")
        api.append(f"  if ({cond}) return {name}::decode(g, Tc, T, imax, C, V, MC, MV, ch, B, early, out);")
    api.append("  return -1000;  // configuration not compiled in\n}")
    api.append("int ref_llr_decode(int DC, int DV, int algo, int N, int M, int E,"
               " const int* sc, const int* dc, const int* tc, const int* sv, const int* dv, const int* tv,"
               " int imax, const double* ch, int B, int early, double* out) {\n"
               "  ref_graph g{N, M, E, sc, dc, tc, sv, dv, tv};")
    for name, dc, dv in table_llr:
        api.append(f"  if (DC == {dc} && DV == {dv}) return {name}::decode(g, algo, imax, ch, B, early, out);")
    api.append("  return -1000;\n}")
    api.append(r"""
// quantize / quantize_LLR launched as in AWGN_Quantizer_BPSK.py:192,218,239 with global size (Nvar, B)
void ref_quantize(int card, const double* x, const double* limits, int Nvar, int B, int* clusters) {
  g_Nvar = Nvar;
  ndrange(Nvar, B, [&] { qnt::quantize(card, x, limits, clusters); });
}
void ref_quantize_llr(int card, const double* x, const double* limits, const double* llr, int Nvar, int B, double* out) {
  g_Nvar = Nvar;
  ndrange(Nvar, B, [&] { qnt::quantize_LLR(card, x, limits, llr, out); });
}
int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}""")
    parts.append("\n".join(api))
    src = os.path.join(OUT, "ref_kernels.cpp")
    with open(src, "w") as fh:
        fh.write("\n".join(parts))
    lib = os.path.join(OUT, "libref_kernels.so")
    cmd = ["g++", "-std=c++17", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-w",
           "-I", OUT, src, "-o", lib]
    if not quiet:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=DEFAULT_REF)
    a = ap.parse_args()
    if not os.path.isdir(a.ref):
        print(f"reference tree {a.ref} not present; keeping any prebuilt oracle/_ref", file=sys.stderr)
        sys.exit(0)
    print(build(a.ref))
