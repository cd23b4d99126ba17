"""Run the UNMODIFIED prebuilt reference binary under the fake-MPI shim.

TEST INFRASTRUCTURE ONLY (oracle/).  Used (a) in the authoring container, where
/root/reference exists, to generate the golden fixtures under tests/golden/
(see oracle/refbin/make_golden.py), and (b) by ``bench.py --impl reference`` /
the ``cpu_baseline`` leg, which run the copy of the binary that
``__graft_entry__.build()`` placed in ``oracle/_ref/`` (git-ignored, travels to
the GPU box).  Nothing under compton2d_b200/ imports this module.

Reference facts this relies on (all `path:line` relative to /root/reference):
  * src_20121026/reader.f:31-35    every input record is format(a80, value)
  * src_20121026/reader.f:157-597  record order of input/input.dat
  * src_20121026/reader.f:608-657  one input/input_JJ_KK.dat per zone (11 records)
  * src_20121026/compton2d.f:11    MPI_INIT is the first statement (fork is legal)
  * src_20121026/commonblock.f     COMMON layout (parsed by `CommonLayout`)
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF_DEFAULT = Path("/root/reference")
REFDIR = REPO / "oracle" / "_ref"            # git-ignored build products
SONAMES = ["libmpi_f90.so.0", "libmpi_f77.so.0", "libmpi.so.0",
           "libopen-rte.so.0", "libopen-pal.so.0"]


# --------------------------------------------------------------------------- build
def build_shim(force: bool = False) -> Path:
    """Compile fakempi.c into oracle/_ref/lib/ under the five sonames the binary needs."""
    lib = REFDIR / "lib"
    lib.mkdir(parents=True, exist_ok=True)
    main = lib / "libmpi_f77.so.0"
    src = HERE / "fakempi.c"
    if force or not main.exists() or main.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-Wl,-soname,libmpi_f77.so.0",
                               "-o", str(main), str(src)])
    empty = lib / "_empty.c"
    empty.write_text("/* empty soname placeholder */\n")
    for so in SONAMES:
        if so == "libmpi_f77.so.0":
            continue
        p = lib / so
        if force or not p.exists():
            subprocess.check_call(["gcc", "-shared", "-fPIC", f"-Wl,-soname,{so}", "-o", str(p), str(empty)])
    return lib


def stage_reference(ref: Path = REF_DEFAULT) -> Path | None:
    """Copy the prebuilt binary and its two shipped input templates into oracle/_ref/
    (binary artefacts only -- never sources).  Returns the staged dir or None."""
    src = ref / "src_20121026"
    if not (src / "compton2d").exists():
        return REFDIR if (REFDIR / "compton2d").exists() else None
    REFDIR.mkdir(parents=True, exist_ok=True)
    for name in ("compton2d", "input.dat", "inputm.dat"):
        dst = REFDIR / name
        if not dst.exists() or dst.stat().st_size != (src / name).stat().st_size:
            shutil.copyfile(src / name, dst)
    os.chmod(REFDIR / "compton2d", 0o755)
    return REFDIR


# --------------------------------------------------------------------------- inputs
def build_plcm(ref: Path = REF_DEFAULT, prog: str = "plcm") -> Path | None:
    """Compile one of the reference's own post-processors (postprocessing/plcm.c: light curves, pspt.c: SEDs; K&R C) from
    the source where it lies into oracle/_ref/<prog>.  Only where the reference tree exists; the binary travels to the
    GPU box."""
    src = ref / "postprocessing" / f"{prog}.c"
    out = REFDIR / prog
    if out.exists() and (not src.exists() or out.stat().st_mtime >= src.stat().st_mtime):
        return out
    if not src.exists():
        return None
    REFDIR.mkdir(parents=True, exist_ok=True)
    subprocess.check_call(["gcc", "-std=gnu89", "-w", "-O1", "-include", str(HERE / "gets_shim.h"), "-o", str(out), str(src), "-lm"])
    return out


def run_plcm(events, answers: str, workdir: Path, infile: str = "p001_evb.dat", prog: str = "plcm") -> dict:
    """Run oracle/_ref/plcm on an event list ([n,7] array, written in the reference's format 105 of imcleak2d.f:181) with
    the given stdin answers (the content of a postprocessing/*_lc.input file).  -> {file name: text} of what it wrote."""
    import numpy as np
    exe = REFDIR / prog
    workdir.mkdir(parents=True, exist_ok=True)
    with open(workdir / infile, "w") as f:
        for r in np.asarray(events, float).reshape(-1, 7):
            f.write(" ".join("%14.7E" % v for v in r) + "\n")
    before = set(os.listdir(workdir))
    # (bytes, not text: plcm echoes file-name buffers it filled with strncpy and never terminated)
    r = subprocess.run(f"ulimit -s unlimited; exec {exe}", shell=True, cwd=workdir, input=answers.encode(), capture_output=True,
                       timeout=600, executable="/bin/bash")
    if r.returncode != 0:
        raise RuntimeError(f"{prog} failed ({r.returncode}): {r.stdout[-500:]!r} {r.stderr[-500:]!r}")
    return {n: (workdir / n).read_text() for n in sorted(set(os.listdir(workdir)) - before)}


def _records(path: Path):
    out = []
    for line in path.read_text().splitlines():
        if not line.strip():
            continue
        out.append((line[:80].ljust(80), line[80:].strip()))
    return out


def _key(label: str) -> str:
    m = re.search(r"([A-Za-z_0-9]+(?:\s*\([^)]*\))?)\s*=\s*$", label)
    return re.sub(r"\s+", "", m.group(1)) if m else label.strip()


def write_inputs(rundir: Path, overrides: dict | None = None, zone: dict | None = None,
                 zone_fn=None, staged: Path | None = None,
                 tbbl=None, l_fname: str | None = None, tbbu=None, u_fname: str | None = None) -> dict:
    """Lay out `rundir` (input/ output/ temp/ rates/) from the shipped templates.

    overrides: {key: value-string} for input.dat, key = the token left of '=' in the
               label (e.g. 'nz', 'nst', 'tstop', 'split1', 'pair_switch').
    zone:      {key: value-string} applied to every per-zone file (keys: tea tna n_e
               ep_switch B_field amxwl gmin gmax p_nth q_turb(j,k) turb_lev(j,k)).
    zone_fn:   optional callable (j, k) -> dict of per-zone overrides.
    tbbl/l_fname: lower-boundary temperature (string) + spectrum file for every k.
    tbbu/u_fname: the same for the upper boundary.
    Returns the effective scalar settings as strings.
    """
    staged = staged or stage_reference() or REFDIR
    overrides = dict(overrides or {})
    recs = _records(staged / "input.dat")
    keys = [_key(l) for l, _ in recs]
    # split template: head (up to ntime), one 4-line boundary group, t0/t1, tail
    i_nt = keys.index("ntime")
    i_t0 = keys.index("t0(1)")
    i_first_bd = i_t0 + 2
    i_spec = keys.index("spec_switch")
    head = recs[: i_nt + 1]
    tt = recs[i_t0: i_t0 + 2]
    grp = recs[i_first_bd: i_first_bd + 4]
    tail = recs[i_spec:]
    vals = {}

    def ov(label, val):
        k = _key(label)
        if k in overrides:
            val = str(overrides[k])
        vals[k] = val
        return label, val

    out = [ov(l, v) for l, v in head] + [ov(l, v) for l, v in tt]
    nr = int(vals["nr"]); nz = int(vals["nz"])
    for k in range(1, nr + 1):
        (l0, v0), (l1, v1), (l2, v2), (l3, v3) = grp
        if tbbu is not None:
            v0 = str(tbbu)
        if u_fname is not None:
            v1 = u_fname
        if tbbl is not None:
            v2 = str(tbbl)
        if l_fname is not None:
            v3 = l_fname
        out += [(l0, v0), (l1, v1), (l2, v2), (l3, v3)]
    out += [ov(l, v) for l, v in tail]
    unused = set(overrides) - set(vals)
    if unused:
        raise KeyError(f"unknown input.dat keys: {sorted(unused)} (have {sorted(set(keys))})")

    for d in ("input", "output", "temp", "rates"):
        (rundir / d).mkdir(parents=True, exist_ok=True)
    (rundir / "input" / "input.dat").write_text("".join(f"{l}{v}\n" for l, v in out))

    zrecs = _records(staged / "inputm.dat")
    for j in range(1, nz + 1):
        for k in range(1, nr + 1):
            z = dict(zone or {})
            if zone_fn:
                z.update(zone_fn(j, k) or {})
            lines = []
            for l, v in zrecs:
                kk = _key(l)
                kk = {"q_turb(j,k)": "q_turb", "turb_lev(j,k)": "turb_lev"}.get(kk, kk)
                lines.append(f"{l}{z.get(kk, v)}\n")
            (rundir / "input" / f"input_{j:02d}_{k:02d}.dat").write_text("".join(lines))
    return vals


# --------------------------------------------------------------------------- run
def run(rundir: Path, trace: bool = True, snapspec: Path | None = None, timeout: float = 3600,
        staged: Path | None = None, cpu: int | None = None) -> subprocess.CompletedProcess:
    lib = build_shim()
    staged = staged or stage_reference() or REFDIR
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = f"{lib}:" + env.get("LD_LIBRARY_PATH", "")
    if trace:
        env["C2D_TRACE"] = str(rundir / "mpitrace")
    if snapspec is not None:
        env["C2D_SNAPSPEC"] = str(snapspec)
        env["C2D_SNAPDIR"] = str(rundir)
    cmd = [str(staged / "compton2d")]
    if cpu is not None:
        cmd = ["taskset", "-c", str(cpu)] + cmd
    with open(rundir / "stdout.txt", "w") as so:
        return subprocess.run(cmd, cwd=rundir, env=env, stdout=so, stderr=subprocess.STDOUT,
                              timeout=timeout, check=False)


def read_trace(rundir: Path, rank: int):
    """-> list of dicts (idx, op, count, dtype, peer, tag, bytes, t)."""
    out = []
    for line in (rundir / f"mpitrace.{rank}").read_text().splitlines():
        if line.startswith("#"):
            if out:
                m = re.match(r"# recv tag (-?\d+) bytes (\d+)", line)
                out[-1]["rtag"] = int(m.group(1)); out[-1]["rbytes"] = int(m.group(2))
            continue
        f = line.split()
        out.append(dict(idx=int(f[0]), op=f[1], count=int(f[2]), dtype=int(f[3]), peer=int(f[4]),
                        tag=int(f[5]), bytes=int(f[6]), t=float(f[7])))
    return out


# --------------------------------------------------------------------------- COMMON layout
class CommonLayout:
    """Byte offsets of every COMMON member, parsed from the reference's own
    commonblock.f + general.pa (read in place, never copied).  ifort lays COMMON
    members out in declaration order; all blocks here are type-homogeneous, so
    no padding questions arise (commonblock.f:118-236)."""

    SIZES = {"integer": 4, "double precision": 8, "real": 4}

    def __init__(self, ref: Path = REF_DEFAULT):
        src = ref / "src_20121026"
        self.params = self._params((src / "general.pa").read_text())
        text = self._join((src / "commonblock.f").read_text())
        self.vars = {}       # name -> (elem_size, dims tuple)
        self.blocks = {}     # block -> [(name, offset, elem_size, dims)]
        for stmt in text:
            s = stmt.strip()
            low = s.lower()
            m = re.match(r"(integer|double precision|real)\s+(.*)", low)
            if m:
                self._decl(self.SIZES[m.group(1)], m.group(2))
                continue
            m = re.match(r"character\s*\*\s*(\d+)\s+(.*)", low)
            if m:
                self._decl(int(m.group(1)), m.group(2))
                continue
            m = re.match(r"common\s*/\s*(\w+)\s*/\s*(.*)", low)
            if m:
                off = 0
                mem = []
                for name in [x.strip() for x in m.group(2).split(",") if x.strip()]:
                    es, dims = self.vars[name]
                    n = 1
                    for d in dims:
                        n *= d
                    mem.append((name, off, es, dims))
                    off += es * n
                self.blocks[m.group(1)] = mem
        self.where = {name: (blk, off, es, dims) for blk, mem in self.blocks.items()
                      for name, off, es, dims in mem}

    @staticmethod
    def _params(text):
        p = {}
        for m in re.finditer(r"parameter\s*\(\s*(\w+)\s*=\s*([^)]+)\)", text, re.I):
            expr = m.group(2).strip().lower().replace("d", "e") if re.search(r"\dd[-+]?\d", m.group(2), re.I) else m.group(2).strip().lower()
            try:
                p[m.group(1).lower()] = eval(expr, {}, p)
            except Exception:
                pass
        return p

    @staticmethod
    def _join(text):
        stmts = []
        for line in text.splitlines():
            if not line.strip() or line[0] in "cC*!" or line.lstrip().startswith("!"):
                continue
            line = line.split("!")[0].rstrip()
            if len(line) > 5 and line[5] not in " 0" and line[:5].strip() == "":
                stmts[-1] += " " + line[6:].strip()
            else:
                stmts.append(line.strip())
        return stmts

    def _decl(self, es, rest):
        depth = 0
        cur = ""
        items = []
        for ch in rest:
            if ch == "(":
                depth += 1
            if ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                items.append(cur); cur = ""
            else:
                cur += ch
        items.append(cur)
        for it in items:
            it = it.strip()
            if not it:
                continue
            m = re.match(r"(\w+)\s*(?:\((.*)\))?$", it)
            name = m.group(1)
            dims = ()
            if m.group(2):
                dims = tuple(int(eval(d.strip(), {}, self.params)) for d in m.group(2).split(","))
            self.vars[name] = (es, dims)

    def ranges(self, name: str, index_lists=None):
        """Byte ranges (symbol, offset, nbytes) covering `name`; index_lists optionally
        restricts each Fortran dimension to range(1, n+1) prefixes: [n1, n2, ...]."""
        blk, off, es, dims = self.where[name.lower()]
        sym = blk + "_"
        if not dims:
            return [(sym, off, es)], ()
        use = list(index_lists or dims)
        # contiguous run along dim 1, iterate over the rest in Fortran order
        out = []
        strides = [1]
        for d in dims[:-1]:
            strides.append(strides[-1] * d)
        import itertools
        rest = [range(u) for u in use[1:]]
        for idx in itertools.product(*reversed(rest)):
            idx = tuple(reversed(idx))
            o = off + es * sum(i * s for i, s in zip(idx, strides[1:]))
            out.append((sym, o, es * use[0]))
        return out, tuple(use)
