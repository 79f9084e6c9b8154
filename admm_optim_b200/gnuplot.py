"""Writer of the drivers' trace files -- the cross-run comparison format of the reference:

  __ADMMStats_step_<k>_.txt          3d_admm.lua:1265-1276 / 2d_admm.lua:1213-1223
  __NewtonStats_step_<k>_.txt        3d_admm.lua:1307-1309 (only with -bNewtonOutput true)
  __NewtonIterations_step_<k>_.txt   3d_admm.lua:1310-1311

The scripts call `gnuplot.write_data(filename, {col1, col2, ...}, false)` of UG4's scripts/util/gnuplot.lua, which is not part of
/root/reference [UPSTREAM-UNVERIFIED].  Restated behaviour: the columns are Lua tables; row r holds col[r] of every column,
values separated by one blank, one row per line; rows run from index 1 up to the first index missing in any column (array part
of a Lua table).  Note what follows for the ADMM table: the scripts fill it at index `admm_steps`, which starts at 0
(3d_admm.lua:874,1265), so the first ADMM iteration of a step is NOT part of the file -- `first_row=0` writes it as well.
Numbers are converted like Lua 5.1's tostring / file:write: "%.14g".
"""
from __future__ import annotations


def lua_number(v) -> str:
    """Lua 5.1 number -> string (LUA_NUMBER_FMT "%.14g"); integers valued floats print without a decimal point, as in Lua."""
    return "%.14g" % float(v)


def write_data(filename, columns, pass_rows=False, mode="w", first_row=1):
    """columns: list of dicts {lua index: value} (or lists, taken as 1-based arrays).  Returns the number of rows written."""
    if pass_rows:
        rows = [list(r.values()) if isinstance(r, dict) else list(r) for r in columns]
    else:
        cols = [c if isinstance(c, dict) else {i + 1: v for i, v in enumerate(c)} for c in columns]
        rows = []
        r = first_row
        while cols and all(r in c for c in cols):
            rows.append([c[r] for c in cols])
            r += 1
    with open(filename, mode) as f:
        for row in rows:
            f.write(" ".join(lua_number(v) for v in row) + " \n")
    return len(rows)


def read_data(filename):
    """Rows of a file written by write_data (or by the reference): list of lists of floats."""
    out = []
    with open(filename) as f:
        for line in f:
            line = line.strip()
            if line and not line.startswith("#"):
                out.append([float(x) for x in line.split()])
    return out
