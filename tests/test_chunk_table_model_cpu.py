"""Host model of the chunk table k_pgs_pipe builds per scene (csrc/pgs_pipe.cu) and of its cursor: levels in order, each cut
into chunks of <= 32 consecutive positions, `base | count << 16`; 32 levels per trip, one per lane, with the lanes' chunk
counts turned into table offsets by an inclusive scan (the shuffles are replaced by array arithmetic here). Checked: the
chunks tile [0, nc) in order, none crosses a level end, none is empty or longer than 32, and the cursor hands out exactly
sweeps x nch entries followed by zeros."""
import numpy as np
import pytest


def build_table(lev_end):
    """lev_end[l], l = 1..ncl (1-based as clev_end; entry 0 = 0)."""
    ncl = len(lev_end) - 1
    tab, nch, carry_end = {}, 0, 0
    for l0 in range(1, ncl + 1, 32):
        lanes = np.arange(32)
        l = l0 + lanes
        e1 = np.array([lev_end[k] if k <= ncl else 0 for k in l])
        e0 = np.concatenate(([carry_end], e1[:-1]))            # __shfl_up by one, lane 0 takes the carry
        carry_end = int(e1[31])
        n = np.where(l <= ncl, (e1 - e0 + 31) >> 5, 0)
        incl = np.cumsum(n)
        off = nch + incl - n
        for lane in range(32):
            for k in range(int(n[lane])):
                tab[int(off[lane]) + k] = int(e0[lane] + 32 * k) | (int(min(32, e1[lane] - e0[lane] - 32 * k)) << 16)
        nch += int(incl[31])
    return [tab[i] for i in range(nch)]


class Cursor:
    def __init__(self, tab, sweeps):
        self.tab, self.nch, self.sweeps, self.k, self.it = tab, len(tab), sweeps, 0, 0
        self.ahead = tab[0] if sweeps > 0 and tab else 0

    def next(self):
        e = self.ahead
        self.k += 1
        if self.k == self.nch:
            self.k, self.it = 0, self.it + 1
        self.ahead = self.tab[self.k] if self.it < self.sweeps else 0
        return e


@pytest.mark.parametrize("seed,nlev", [(1, 1), (2, 31), (3, 32), (4, 33), (5, 130), (6, 700)])
def test_table_tiles_the_stream_level_by_level(seed, nlev):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(1, 90, nlev)                          # rows per level: some need three chunks
    lev_end = np.concatenate(([0], np.cumsum(sizes)))
    tab = build_table(lev_end)
    pos, lev = 0, 1
    for e in tab:
        base, cnt = e & 0xFFFF, (e >> 16) & 0x3F
        assert base == pos and 1 <= cnt <= 32
        assert base + cnt <= lev_end[lev]                      # inside its level
        pos += cnt
        if pos == lev_end[lev]:
            lev += 1
    assert pos == lev_end[-1] and lev == nlev + 1
    assert len(tab) == int(((sizes + 31) // 32).sum())


@pytest.mark.parametrize("sweeps", [0, 1, 10])
def test_cursor_hands_out_every_sweep_then_zeros(sweeps):
    tab = build_table(np.concatenate(([0], np.cumsum([5, 40, 32, 1]))))
    c = Cursor(tab, sweeps)
    got = [c.next() for _ in range(sweeps * len(tab) + 3)]
    assert got[:sweeps * len(tab)] == tab * sweeps and got[sweeps * len(tab):] == [0, 0, 0]
