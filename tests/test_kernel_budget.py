"""Static budget of the persistent update kernels, read from the ptxas summary the build leaves next to the sources
(csrc/Makefile: -Xptxas -v).  Two regressions cost 15 % of the update's time in round 2 and are invisible to every
functional test: kernel parameters taken by value and passed on by reference (every thread then starts by copying
~750 bytes of parameters to its stack and every callee reads them back from local memory; they are `__grid_constant__`
now), and a search function whose straight-line code spills a few hundred bytes per thread."""
import re
from pathlib import Path

LOG = Path(__file__).resolve().parents[1] / "agi_lidar_slam_b200" / "csrc" / "lio_pass.ptxas.log"


def _entries():
    txt = LOG.read_text()
    out = {}
    for m in re.finditer(r"Compiling entry function '(\w+)' for 'sm_100a'\n.*?Function properties for \1\n\s+(\d+) bytes stack frame, "
                         r"(\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt, re.S):
        out[m.group(1)] = tuple(int(m.group(k)) for k in (2, 3, 4, 5))
    return txt, out


def test_update_kernels_keep_their_parameters_out_of_local_memory():
    assert LOG.exists(), "build first: make -C agi_lidar_slam_b200/csrc"
    _, ent = _entries()
    upd = {k: v for k, v in ent.items() if "update_kernel" in k or "pass_kernel" in k}
    assert len(upd) >= 6, sorted(ent)
    for name, (stack, st, ld, regs) in upd.items():
        assert stack <= 384, (name, stack)  # 752 bytes with the parameter copies
        assert st <= 64 and ld <= 64, (name, st, ld)
        assert regs <= 128, (name, regs)  # one 512-thread block per SM


def test_search_tiles_do_not_spill_much():
    txt, _ = _entries()
    seen = 0
    for m in re.finditer(r"Function properties for (\w*search_tile\w*)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                         r"(\d+) bytes spill loads", txt):
        seen += 1
        assert int(m.group(2)) == 0 and int(m.group(3)) <= 128 and int(m.group(4)) <= 128, m.group(0)
    assert seen >= 6
