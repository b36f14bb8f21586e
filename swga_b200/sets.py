"""Candidate-set enumeration -- the non-data-parallel remainder, kept on the host.

Same seam as swga/sets.py:9-71: `find(...)` returns a generator over the stdout lines of the
`set_finder` program (cliquer with swga's front end, ext/cliquer/set_finder.c), one candidate set per
line, `"id,id,... <bg_len/sum(weights) as %f>"` (set_finder.c:427-444); closing the generator kills
the process group.  The binary is compiled unchanged from the reference's sources by
`swga_b200.build.build_set_finder()` into swga_b200/bin/ (or named by $SWGA_SET_FINDER).
"""
import os
import signal
import subprocess
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def set_finder():
    """Path of the set_finder executable (swga/utils.py:set_finder in the reference)."""
    cand = [os.environ.get("SWGA_SET_FINDER"), os.path.join(HERE, "bin", "set_finder")]
    for c in cand:
        if c and os.access(c, os.X_OK):
            return c
    from . import build
    p = build.build_set_finder()
    if p is None:
        raise IOError("set_finder not found: set $SWGA_SET_FINDER or build it from the reference's ext/cliquer "
                      "(python -m swga_b200.build)")
    return p


def find(**kwargs):
    """Find sets using set_finder.

    :param min_bg_bind_dist: min avg bg binding distance between primers
    :param bg_length: bg genome length
    :param min_size: smallest allowable set size
    :param max_size: largest allowable set size
    :param graph_fp: compatibility graph file (DIMACS)
    :param workers: <= 1 searches in weighted-coloring order; > 1 runs that many searches from random
                    vertex orders and interleaves their output round-robin (non-deterministic).
    """
    workers = kwargs.get("workers", 1)
    if workers <= 1:
        return _find_sets(**{k: v for k, v in kwargs.items() if k != "workers"})
    return _mp_find_sets(**kwargs)


def _find_sets(min_bg_bind_dist, bg_length, min_size, max_size, graph_fp, vertex_ordering="weighted-coloring"):
    assert vertex_ordering in ("weighted-coloring", "random")
    cmd = [set_finder(), "-q", "-q", "--bg_freq", str(min_bg_bind_dist), "--bg_len", str(bg_length),
           "--min", str(min_size), "--max", str(max_size), "--unweighted", "--all",
           "--reorder", vertex_ordering, str(graph_fp)]
    process = subprocess.Popen(cmd, stdout=subprocess.PIPE, preexec_fn=os.setsid, bufsize=0)
    try:
        for line in iter(process.stdout.readline, b""):
            yield line.decode("ascii", "replace")
    finally:
        time.sleep(0.1)
        if process.poll() is None:
            os.killpg(process.pid, signal.SIGKILL)
        process.stdout.close()
        process.wait()


def find_blocks(min_bg_bind_dist, bg_length, min_size, max_size, graph_fp, vertex_ordering="weighted-coloring",
                block_bytes=1 << 20):
    """The same set_finder stream as `find(workers=1)`, handed over in raw blocks of stdout (bytes, not aligned
    to lines) for score.parse_set_lines: no per-line Python between cliquer and the scoring kernel."""
    cmd = [set_finder(), "-q", "-q", "--bg_freq", str(min_bg_bind_dist), "--bg_len", str(bg_length),
           "--min", str(min_size), "--max", str(max_size), "--unweighted", "--all",
           "--reorder", vertex_ordering, str(graph_fp)]
    process = subprocess.Popen(cmd, stdout=subprocess.PIPE, preexec_fn=os.setsid, bufsize=0)
    try:
        while True:
            chunk = process.stdout.read(block_bytes)
            if not chunk:
                break
            yield chunk
    finally:
        time.sleep(0.1)
        if process.poll() is None:
            os.killpg(process.pid, signal.SIGKILL)
        process.stdout.close()
        process.wait()


def _mp_find_sets(workers, **kwargs):
    procs = [_find_sets(vertex_ordering="random", **kwargs) for _ in range(workers)]
    try:
        while True:
            for p in procs:
                try:
                    yield next(p)
                except StopIteration:
                    return
    finally:
        for p in procs:
            p.close()
