"""A second, independent restatement of the reference's cell list and neighbour loop — plain Python, line by line,
with the reference's own data structures (growable zero-terminated index vectors per cell) instead of the oracle's CSR
arrays.  Slow by design (small inputs only); used by tests/test_oracle_pins.py to cross-check the C++ oracle on random
inputs, including the corners the reference's own tests never touch: narrow domains (double visits), row wrap-around of
the linear stencil, particles on cell and domain faces, duplicates, NaN/Inf positions, repeated rebuilds.

Reference: src/structs.jl:57-106 (constructor, find_key), src/core.jl:8-112 (dist, add_index!, create_cell_list!,
_apply_binary!), src/geometry.jl:24-30 (is_inside(Box))."""
import math


class LiteralSystem:
    def __init__(self, lo, hi, h):  # structs.jl:57-91
        assert h > 0.0
        self.h = h
        self.lo, self.hi = tuple(lo), tuple(hi)
        self.key_phase = [int(math.floor(v / h)) for v in lo]
        self.key_lim = [int(math.floor(v / h)) - p + 1 for v, p in zip(hi, self.key_phase)]
        self.key_max = self.key_lim[0] * self.key_lim[1] * self.key_lim[2]
        self.key_diff = []
        if self.key_lim[2] == 1:
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    self.key_diff.append(di + self.key_lim[0] * dj)
        else:
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    for dk in (-1, 0, 1):
                        self.key_diff.append(di + self.key_lim[0] * (dj + self.key_lim[1] * dk))
        self.particles = []                                   # each particle: a dict with "x" and "tag"
        self.cell_list = [[0] for _ in range(self.key_max)]   # Cell.entries, structs.jl:22-31
        self.removal_cell = [0]

    def find_key(self, x):  # structs.jl:97-106
        try:
            i = 1 + int(math.floor(x[0] / self.h)) - self.key_phase[0]
            j = 1 + int(math.floor(x[1] / self.h)) - self.key_phase[1]
            k = 1 + int(math.floor(x[2] / self.h)) - self.key_phase[2]
            return i + self.key_lim[0] * (j - 1) + self.key_lim[0] * self.key_lim[1] * (k - 1)
        except (ValueError, OverflowError):  # Int64(x) fails when x is nan or infinity
            return -1

    def is_inside(self, x):  # geometry.jl:24-30 (comparisons with NaN are false)
        return (self.lo[0] <= x[0] <= self.hi[0] and self.lo[1] <= x[1] <= self.hi[1] and self.lo[2] <= x[2] <= self.hi[2])

    @staticmethod
    def add_index(entries, i):  # core.jl:13-41
        ind = None
        for t, v in enumerate(entries):      # find_vacation!
            if v == 0:
                ind = t
                break
        if ind is None:
            entries.append(0)
            ind = len(entries) - 1
        entries[ind] = i
        while ind > 0 and entries[ind - 1] < entries[ind]:
            entries[ind], entries[ind - 1] = entries[ind - 1], entries[ind]
            ind -= 1

    def create_cell_list(self):  # core.jl:51-90
        for cell in self.cell_list:
            for k in range(len(cell)):
                cell[k] = 0
        for k in range(len(self.removal_cell)):
            self.removal_cell[k] = 0
        for i in range(1, len(self.particles) + 1):
            if not self.is_inside(self.particles[i - 1]["x"]):
                self.add_index(self.removal_cell, i)
        i = 1
        while i <= len(self.removal_cell) and self.removal_cell[i - 1] != 0:
            self.particles[self.removal_cell[i - 1] - 1] = self.particles[len(self.particles) + 1 - i - 1]
            i += 1
        if i > 1:
            del self.particles[len(self.particles) + 1 - i:]
        for i in range(1, len(self.particles) + 1):
            key = self.find_key(self.particles[i - 1]["x"])
            self.add_index(self.cell_list[key - 1], i)

    def neighbours(self, ip):  # _apply_binary!, core.jl:94-112: the (j) of every action!(p, q, r) call, in order
        p = self.particles[ip - 1]
        key = self.find_key(p["x"])
        out = []
        for dkey in self.key_diff:
            nk = key + dkey
            if 1 <= nk <= self.key_max:
                for j in self.cell_list[nk - 1]:
                    if j == 0:
                        break
                    q = self.particles[j - 1]
                    d = [p["x"][c] - q["x"][c] for c in range(3)]
                    r = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])   # dist, core.jl:8-10 / algebra.jl:49-60
                    if r > self.h or q is p:
                        continue
                    out.append(j)
        return out
