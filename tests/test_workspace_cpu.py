"""The workspace database (swga_b200/workspace.py) against the reference's model definitions
(swga/workspace.py:106-247) and its own tests (swga/tests/test_workspace.py:14-18,
swga/tests/test_primers.py): tables, columns, keys, defaults, Set.add semantics, version check."""
import math
import sqlite3

import pytest

from swga_b200 import workspace


@pytest.fixture
def ws(tmp_path):
    w = workspace.Workspace(str(tmp_path / "primers.db"))
    w.create_tables()
    w.metadata = dict(version=workspace.VERSION, fg_file="fg.fa", bg_file="bg.fa", ex_file=None, fg_length=1000,
                      bg_length=20000)
    yield w
    w.close()


def test_tables_and_columns(ws):
    # swga/tests/test_workspace.py:14-18 expects the four tables of the models
    assert set(ws.tables()) == {"primer", "set", "set_primer_through", "_metadata"}
    cols = {r["name"]: r for r in ws.db.execute('PRAGMA table_info("primer")')}
    assert tuple(cols) == workspace.PRIMER_COLUMNS
    assert cols["seq"]["pk"] == 1 and cols["_id"]["notnull"] == 0 and cols["tm"]["notnull"] == 0
    assert [r["name"] for r in ws.db.execute('PRAGMA table_info("set")')] == list(workspace.SET_COLUMNS)
    m = ws.metadata
    assert m["fg_length"] == 1000 and m["bg_length"] == 20000 and m["db_name"].endswith("primers.db")


def test_primer_defaults_and_primary_key(ws):
    ws.add_primers([dict(seq="AAGT", fg_freq=3, bg_freq=0, ratio=float("inf"))], add_revcomp=True)
    ps = {p.seq: p for p in ws.primers()}
    assert set(ps) == {"AAGT", "ACTT"}                                   # primers.py:77-82
    p = ps["ACTT"]
    assert (p.fg_freq, p.bg_freq, p.active, p._id, p.tm, p.gini, p.locations) == (3, 0, False, None, None, None, None)
    assert math.isinf(p.ratio)
    with pytest.raises(sqlite3.IntegrityError):                          # seq is the primary key
        ws.add_primers([dict(seq="AAGT", fg_freq=1, bg_freq=1, ratio=1.0)], add_revcomp=False)
    with pytest.raises(sqlite3.IntegrityError):                          # a palindrome collides with its own revcomp
        ws.add_primers([dict(seq="ACGT", fg_freq=1, bg_freq=1, ratio=1.0)], add_revcomp=True)
    assert ws.count_primers() == 2                                       # the failed inserts rolled back


def test_update_and_locations_round_trip(ws):
    ws.add_primers([dict(seq="AAGT", fg_freq=3, bg_freq=2, ratio=1.5)], add_revcomp=False)
    p = ws.primers()[0]
    p.locations, p.gini, p.active, p._id = {"chr1": [1, 7], "chr2": []}, 0.25, True, 4
    ws.update_primers([p])
    q = ws.primers('"active" = 1')[0]
    assert (q.locations, q.gini, q.active, q._id) == ({"chr1": [1, 7], "chr2": []}, 0.25, True, 4)
    ws.deactivate_all()
    assert ws.primers('"active" = 1') == []
    ws.reset_ids()
    assert ws.primers()[0]._id == -1


def test_set_add_is_idempotent_per_primer_set(ws):
    ws.add_primers([dict(seq=s, fg_freq=1, bg_freq=1, ratio=1.0) for s in ("AAGT", "CCGA", "TTGA")], add_revcomp=False)
    var = dict(set_size=2, bg_dist_mean=10.0, fg_dist_std=1.0, fg_dist_gini=0.5, fg_max_dist=7, fg_dist_mean=3.5)
    assert ws.add_set(1, ["AAGT", "CCGA"], 0.175, "expr", **var) == (1, True)
    assert ws.add_set(2, ["CCGA", "AAGT"], 0.175, "expr", **var) == (1, False)      # same set, any order
    assert ws.add_set(2, ["CCGA", "TTGA"], 0.5, "expr", **dict(var, fg_dist_std=float("nan"))) == (2, True)
    with pytest.raises(ValueError):
        ws.add_set(3, [], 0.0, "expr")
    rows = ws.sets()
    assert [r["_id"] for r in rows] == [1, 2] and rows[0]["primers"] == ["AAGT", "CCGA"]
    assert rows[0]["fg_max_dist"] == 7 and rows[0]["scoring_fn"] == "expr" and math.isnan(rows[1]["fg_dist_std"])
    assert ws.min_set_id() == 1
    ws.reset_sets()
    assert ws.count_sets() == 0 and ws.count_primers() == 3
    ws.reset_primers()
    assert ws.count_primers() == 0 and set(ws.tables()) == {"primer", "set", "set_primer_through", "_metadata"}


def test_version_check(tmp_path):
    path = str(tmp_path / "w.db")
    with workspace.Workspace(path) as w:
        w.create_tables()
        w.metadata = dict(version="0.3.9", fg_file="a", bg_file="b", ex_file=None, fg_length=1, bg_length=1)
    with pytest.raises(workspace.WorkspaceError):
        workspace.connection(path)
    with workspace.Workspace(path) as w:
        w.metadata = dict(version="0.4.0", fg_file="a", bg_file="b", ex_file=None, fg_length=1, bg_length=1)
    workspace.connection(path).close()                                   # same major.minor: accepted
    with pytest.raises(workspace.WorkspaceError):
        workspace.connection(str(tmp_path / "missing.db"))
