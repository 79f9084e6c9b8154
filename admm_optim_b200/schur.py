"""Host-side dense m x m algebra with the semantics of the reference's vendored lua-matrix
(lua-matrix/matrix.lua), used by the drivers for the Schur complement (3d_admm.lua:634-645, 1063-1078).
m = dim+1 <= 4, so this stays on the host exactly like in the reference.

`invert` = Gauss-Jordan on [S | I] (matrix.lua:513-534, dogauss :450-507) whose pivot rule picks the row
with the SMALLEST non-zero |entry| (pivotOk, :422-442) -- restated because it decides the last digits of
DeltaLambda.  `mul` accumulates left to right (:223-237).
"""
from __future__ import annotations

import math


class Matrix(list):
    def __init__(self, rows, cols=None, value=0.0):
        if isinstance(rows, int):
            super().__init__([[value] * (cols if cols is not None else rows) for _ in range(rows)])
        else:
            super().__init__([list(r) for r in rows])

    @property
    def nrows(self):
        return len(self)

    @property
    def ncols(self):
        return len(self[0])

    def copy(self):
        return Matrix(self)

    def mul(self, other):
        assert self.ncols == len(other), "matrix size mismatch"
        out = Matrix(self.nrows, len(other[0]))
        for i in range(self.nrows):
            for j in range(len(other[0])):
                num = self[i][0] * other[0][j]
                for n in range(1, self.ncols):
                    num = num + self[i][n] * other[n][j]
                out[i][j] = num
        return out

    def sub(self, other):
        return Matrix([[self[i][j] - other[i][j] for j in range(self.ncols)] for i in range(self.nrows)])

    def invert(self):
        """Returns the inverse, or None when the matrix is singular (lua: nil, rank)."""
        n = self.nrows
        assert n == self.ncols, "matrix not square"
        mtx = [list(map(float, self[i])) + [1.0 if i == j else 0.0 for j in range(n)] for i in range(n)]
        columns = 2 * n
        for j in range(n):                      # stairs left -> right
            i_min, norm_min = None, math.inf
            for i in range(j, n):               # pivotOk: smallest non-zero magnitude
                norm = abs(mtx[i][j])
                if norm > 0 and norm < norm_min:
                    i_min, norm_min = i, norm
            if i_min is None:
                return None
            if i_min != j:
                mtx[j], mtx[i_min] = mtx[i_min], mtx[j]
            for i in range(j + 1, n):
                if mtx[i][j] != 0:
                    factor = mtx[i][j] / mtx[j][j]
                    mtx[i][j] = 0.0
                    for c in range(j + 1, columns):
                        mtx[i][c] = mtx[i][c] - factor * mtx[j][c]
        for j in range(n - 1, -1, -1):          # stairs right <- left
            div = mtx[j][j]
            for c in range(j + 1, columns):
                mtx[j][c] = mtx[j][c] / div
            for i in range(j - 1, -1, -1):
                if mtx[i][j] != 0:
                    factor = mtx[i][j]
                    for c in range(j + 1, columns):
                        mtx[i][c] = mtx[i][c] - factor * mtx[j][c]
                    mtx[i][j] = 0.0
            mtx[j][j] = 1.0
        return Matrix([row[n:] for row in mtx])

    def print(self):
        for row in self:
            print("\t".join(repr(v) for v in row))
