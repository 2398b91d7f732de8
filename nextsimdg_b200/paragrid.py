"""The dynamics state in the reference's ParaGrid restart layout.

`ParaGridIO::dumpModelState` (core/src/ParaGridIO.cpp:272-348) writes a netCDF-4 file with the groups `structure`,
`metadata` (time + configuration) and `data`; the data group carries the dimensions of
core/src/discontinuousgalerkin/ModelArrayDetails.cpp:30-43 (xdim, ydim, zdim, xvertex, yvertex, x_cg, y_cg, dg_comp,
dgstress_comp, ncoords) and one double variable per restart field, dimensions slowest first with the component index
last -- `hice(ydim, xdim, dg_comp)`, `u(ydim, xdim)`, `coords(yvertex, xvertex, ncoords)` -- and a `missing_value`
attribute (ParaGridIO.cpp:339-341).  The same layout is what the reference's init-file writers produce
(run/make_init_base.py:112-186).

netCDF is absent from this image, so the container here is the CDL TEXT of such a file (`ncgen -o restart.nc
restart.cdl` turns it into the binary the model reads; `ncdump` gives the text back): `to_cdl` writes it, `from_cdl`
reads it.  Beyond the fields the reference checkpoints for the dynamics (hice, cice, damage as full DG fields and the
cell-mean u, v: BBMDynamics.cpp:104-132, IDynamics.hpp:46-53), the CG velocity, the three DG stresses and (BBM) the
running-mean velocity are stored as extra variables with a `dynamics_` prefix: the reference restarts from zero stress
and the DG0 velocity (CGDynamicsKernel.cpp:57 TODO); with them `restore` resumes bit-reproducibly.
"""
from __future__ import annotations

import re

import numpy as np

MISSING = 1.7e38  # MissingData::value, core/src/include/MissingData.hpp:17
STRUCTURE = "parametric_rectangular"  # ParametricGrid::structureName

# variable -> dimensions (slowest first), ParaGridIO.cpp:298-311 + ModelArrayDetails.cpp:45-95
DIMS = {
    "mask": ("ydim", "xdim"), "coords": ("yvertex", "xvertex", "ncoords"), "u": ("ydim", "xdim"), "v": ("ydim", "xdim"),
    "hice": ("ydim", "xdim", "dg_comp"), "cice": ("ydim", "xdim", "dg_comp"), "damage": ("ydim", "xdim", "dg_comp"),
    "dynamics_cg_u": ("y_cg", "x_cg"), "dynamics_cg_v": ("y_cg", "x_cg"), "dynamics_avg_u": ("y_cg", "x_cg"),
    "dynamics_avg_v": ("y_cg", "x_cg"), "dynamics_s11": ("ydim", "xdim", "dgstress_comp"),
    "dynamics_s12": ("ydim", "xdim", "dgstress_comp"), "dynamics_s22": ("ydim", "xdim", "dgstress_comp"),
}
_INTERNAL = {"dynamics_cg_u": "cg_u", "dynamics_cg_v": "cg_v", "dynamics_avg_u": "avgU", "dynamics_avg_v": "avgV",
             "dynamics_s11": "s11", "dynamics_s12": "s12", "dynamics_s22": "s22"}


def dimensions(nx: int, ny: int, dgadv: int = 6, cgdegree: int = 2) -> dict:
    dgs = 8 if cgdegree == 2 else 3  # CG2DGSTRESS, NextsimDynamics.hpp:42-60
    return {"xdim": nx, "ydim": ny, "zdim": 1, "xvertex": nx + 1, "yvertex": ny + 1, "x_cg": cgdegree * nx + 1,
            "y_cg": cgdegree * ny + 1, "dg_comp": dgadv, "dgstress_comp": dgs, "ncoords": 2}


def restart_state(dyn, ms: dict | None = None) -> dict:
    """name -> array in the ParaGrid shapes, from a CUDADynamicsBase handle (and optionally the mesh from `ms`)."""
    nx, ny = dyn.nx, dyn.ny
    out = {}
    if ms is not None:
        out["mask"] = np.asarray(ms["mask"], dtype=np.float64).reshape(ny, nx)
        out["coords"] = np.asarray(ms["coords"], dtype=np.float64).reshape(ny + 1, nx + 1, 2)
    out["hice"], out["cice"] = dyn.getDGData("hice"), dyn.getDGData("cice")
    if dyn.usesDamage():
        out["damage"] = dyn.getDGData("damage")
    out["u"], out["v"] = np.asarray(dyn.uice, dtype=np.float64), np.asarray(dyn.vice, dtype=np.float64)
    d = dimensions(nx, ny, dyn.dgadv, dyn.cgdegree)
    for name, internal in _INTERNAL.items():
        if internal.startswith("avg") and not dyn.usesDamage():
            continue
        out[name] = dyn.internal(internal).reshape([d[k] for k in DIMS[name]])
    return out


def restore(dyn, state: dict):
    """Put a `restart_state` back into a handle whose mesh is set (bit-reproducible resume)."""
    for name in ("hice", "cice", "damage"):
        if name in state and (name != "damage" or dyn.usesDamage()):
            dyn._set(name, state[name])
    dyn.uice, dyn.vice = np.array(state["u"], dtype=np.float64), np.array(state["v"], dtype=np.float64)
    for name, internal in _INTERNAL.items():
        if name in state:
            dyn.set_internal(internal, state[name])


def to_cdl(state: dict, nx: int, ny: int, dgadv: int = 6, cgdegree: int = 2, time_unix: int = 0,
           time_formatted: str = "1970-01-01T00:00:00Z", name: str = "restart") -> str:
    """CDL text of the restart file ParaGridIO::dumpModelState would write for these fields."""
    d = dimensions(nx, ny, dgadv, cgdegree)
    lines = [f"netcdf {name} {{", "", "group: structure {", "", "  // group attributes:", f'\t\t:type = "{STRUCTURE}" ;',
             "  } // group structure", "", "group: metadata {", "", "  // group attributes:", f'\t\t:type = "{STRUCTURE}" ;', "",
             "  group: time {", "    variables:", "\tint64 time ;", '\t\ttime:units = "seconds since 1970-01-01T00:00:00Z" ;',
             '\t\ttime:format = "%Y-%m-%dT%H:%M:%SZ" ;', f'\t\ttime:formatted = "{time_formatted}" ;', "    data:", "",
             f"     time = {int(time_unix)} ;", "    } // group time", "", "  group: configuration {", "    } // group configuration",
             "  } // group metadata", "", "group: data {", "  dimensions:"]
    lines += [f"\t{k} = {v} ;" for k, v in d.items()]
    lines.append("  variables:")
    for k, a in state.items():
        dims = DIMS[k]
        want = tuple(d[x] for x in dims)
        if tuple(np.shape(a)) != want:
            raise ValueError(f"{k}: shape {np.shape(a)} does not match the ParaGrid layout {dims} = {want}")
        lines += [f"\tdouble {k}({', '.join(dims)}) ;", f"\t\t{k}:missing_value = {MISSING!r} ;"]
    lines.append("  data:")
    for k, a in state.items():
        flat = np.asarray(a, dtype=np.float64).ravel()
        lines += ["", f"   {k} = " + ", ".join(repr(float(x)) for x in flat) + " ;"]  # repr round-trips every double
    lines += ["  } // group data", "}", ""]
    return "\n".join(lines)


def from_cdl(text: str) -> tuple[dict, dict]:
    """(state, dimensions) of the data group of a restart file in CDL text form."""
    m = re.search(r"group:\s*data\s*\{(.*)\}\s*//\s*group data", text, re.S)
    if not m:
        raise ValueError("no data group in the CDL text")
    body = m.group(1)
    dims_txt, rest = body.split("variables:", 1)
    var_txt, data_txt = rest.split("data:", 1)
    dims = {k: int(v) for k, v in re.findall(r"(\w+)\s*=\s*(\d+)\s*;", dims_txt)}
    shapes = {k: tuple(dims[x.strip()] for x in dd.split(",")) for k, dd in re.findall(r"double\s+(\w+)\(([^)]*)\)\s*;", var_txt)}
    state = {}
    for k, vals in re.findall(r"(\w+)\s*=\s*([^;]*);", data_txt):
        if k in shapes:
            state[k] = np.array([float(x) for x in vals.replace("\n", " ").split(",")], dtype=np.float64).reshape(shapes[k])
    missing = set(shapes) - set(state)
    if missing:
        raise ValueError(f"variables without data: {sorted(missing)}")
    return state, dims
