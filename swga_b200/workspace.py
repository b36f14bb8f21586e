"""The swga workspace database on stdlib sqlite3.

The reference declares its tables with peewee 2.x models (swga/workspace.py:106-247); peewee is not
part of this path, so the DDL below restates what those models create: tables `primer`, `set`, the
many-to-many through table of `Set.primers` and `_metadata`, with the reference's column names, types,
defaults and keys, so that a workspace written here opens in the reference and vice versa.  (peewee
names the through table `<lhs>_<rhs>_through` with foreign-key columns `<model>_id`; that part is
inferred from peewee 2.x's ManyToManyField, the source is not vendored in the reference.)

`Set._hash` is `hash(frozenset(seqs))` of a Python-2 process in the reference (workspace.py:228):
unstable across interpreters, so here it is a stable 63-bit FNV-1a of the sorted sequences -- it only
has to be equal for equal primer sets inside one workspace.
"""
import json
import os
import sqlite3

VERSION = "0.4.4"                                   # swga/__init__.py: workspaces are valid for major.minor
DEFAULT_DB_FNAME = "primers.db"

PRIMER_COLUMNS = ("_id", "seq", "fg_freq", "bg_freq", "ratio", "tm", "_locations", "gini", "active")
SET_COLUMNS = ("_id", "_hash", "score", "set_size", "bg_dist_mean", "fg_dist_std", "fg_dist_gini", "fg_max_dist",
               "fg_dist_mean", "scoring_fn")
SET_VARIABLES = ("set_size", "bg_dist_mean", "fg_dist_std", "fg_dist_gini", "fg_max_dist", "fg_dist_mean")

_DDL = {
    "primer": """CREATE TABLE IF NOT EXISTS "primer" (
        "_id" INTEGER, "seq" VARCHAR(255) NOT NULL PRIMARY KEY, "fg_freq" INTEGER NOT NULL DEFAULT 0,
        "bg_freq" INTEGER NOT NULL DEFAULT 0, "ratio" REAL NOT NULL DEFAULT 0.0, "tm" REAL, "_locations" TEXT,
        "gini" REAL, "active" SMALLINT NOT NULL DEFAULT 0)""",
    "set": """CREATE TABLE IF NOT EXISTS "set" (
        "_id" INTEGER NOT NULL PRIMARY KEY, "_hash" INTEGER, "score" REAL NOT NULL, "set_size" INTEGER,
        "bg_dist_mean" REAL, "fg_dist_std" REAL, "fg_dist_gini" REAL, "fg_max_dist" INTEGER, "fg_dist_mean" REAL,
        "scoring_fn" TEXT)""",
    "set_hash": 'CREATE UNIQUE INDEX IF NOT EXISTS "set__hash" ON "set" ("_hash")',
    "through": """CREATE TABLE IF NOT EXISTS "set_primer_through" (
        "id" INTEGER NOT NULL PRIMARY KEY, "set_id" INTEGER NOT NULL REFERENCES "set" ("_id"),
        "primer_id" VARCHAR(255) NOT NULL REFERENCES "primer" ("seq"))""",
    "through_idx": 'CREATE UNIQUE INDEX IF NOT EXISTS "set_primer_through_set_id_primer_id" '
                   'ON "set_primer_through" ("set_id", "primer_id")',
    "_metadata": """CREATE TABLE IF NOT EXISTS "_metadata" (
        "id" INTEGER NOT NULL PRIMARY KEY, "db_name" TEXT NOT NULL, "version" TEXT, "fg_file" TEXT, "bg_file" TEXT,
        "ex_file" TEXT, "fg_length" INTEGER, "bg_length" INTEGER)""",
}


class WorkspaceError(RuntimeError):
    pass


def set_hash(seqs):
    """Stable stand-in for the reference's hash(frozenset(seqs)) (workspace.py:228)."""
    h = 0xcbf29ce484222325
    for s in sorted(set(seqs)):
        for b in s.encode("ascii") + b"\0":
            h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h >> 1                                   # sqlite INTEGER is signed 64-bit


class Workspace(object):
    """One workspace database (swga/workspace.py:23-77, 250-262)."""

    def __init__(self, db_name=DEFAULT_DB_FNAME):
        self.db_name = db_name
        self.db = sqlite3.connect(db_name)
        self.db.row_factory = sqlite3.Row

    # -- context manager: workspace.connection(db_name) (workspace.py:250-262)
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self.db is not None:
            self.db.commit()
            self.db.close()
            self.db = None

    # -- tables
    def create_tables(self):
        for key in ("primer", "set", "set_hash", "through", "through_idx", "_metadata"):
            self.db.execute(_DDL[key])
        self.db.commit()

    def tables(self):
        return [r[0] for r in self.db.execute("SELECT name FROM sqlite_master WHERE type='table' ORDER BY name")]

    def reset_sets(self):                           # workspace.py:42-44
        self.db.execute('DROP TABLE IF EXISTS "set_primer_through"')
        self.db.execute('DROP TABLE IF EXISTS "set"')
        for key in ("set", "set_hash", "through", "through_idx"):
            self.db.execute(_DDL[key])
        self.db.commit()

    def reset_primers(self):                        # workspace.py:46-48
        self.db.execute('DROP TABLE IF EXISTS "set_primer_through"')
        self.db.execute('DROP TABLE IF EXISTS "set"')
        self.db.execute('DROP TABLE IF EXISTS "primer"')
        self.create_tables()

    # -- metadata (workspace.py:26-36)
    @property
    def metadata(self):
        r = self.db.execute('SELECT * FROM "_metadata" LIMIT 1').fetchone()
        if r is None:
            raise WorkspaceError("workspace has no metadata: run init first")
        return {k: r[k] for k in r.keys() if k != "id"}

    @metadata.setter
    def metadata(self, values):
        self.db.execute('DELETE FROM "_metadata"')
        cols = ["db_name"] + list(values)
        self.db.execute('INSERT INTO "_metadata" (%s) VALUES (%s)' % (", ".join('"%s"' % c for c in cols),
                                                                     ", ".join("?" * len(cols))),
                        [self.db_name] + [values[c] for c in values])
        self.db.commit()

    def check_version(self, version=VERSION):
        """workspace.py:50-77: the workspace's major.minor must equal the running version's."""
        try:
            db_ver = self.metadata["version"]
        except (sqlite3.OperationalError, WorkspaceError):
            db_ver = "<NA>"
        if str(db_ver).split(".")[:2] != str(version).split(".")[:2]:
            raise WorkspaceError("This workspace was created with a different version of swga.\n"
                                 "  Workspace version: {}\n  swga version:      {}".format(db_ver, version))

    # -- primers (swga/primers.py:71-87 and the Primer model)
    def add_primers(self, primer_dicts, add_revcomp=True):
        """Primers.add: rows {'seq','fg_freq','bg_freq','ratio'}; with add_revcomp the reverse complement of
        each is added with the same numbers.  `seq` is the primary key: a duplicate raises, as in the
        reference (palindromes are therefore rejected when add_revcomp is set, count.py:112)."""
        from .locate import revcomp
        rows = list(primer_dicts)
        if add_revcomp:
            rows = rows + [dict(r, seq=revcomp(r["seq"])) for r in rows]
        with self.db:
            self.db.executemany('INSERT INTO "primer" ("seq", "fg_freq", "bg_freq", "ratio") VALUES (?, ?, ?, ?)',
                                [(r["seq"], int(r["fg_freq"]), int(r["bg_freq"]), _sql_float(r["ratio"])) for r in rows])
        return len(rows)

    def count_primers(self):
        return self.db.execute('SELECT COUNT(*) FROM "primer"').fetchone()[0]

    def primers(self, where="", args=(), order_by=""):
        """Rows of `primer` as pipeline.Primer records (locations decoded from the JSON column)."""
        from .pipeline import Primer
        q = 'SELECT * FROM "primer"' + (" WHERE " + where if where else "") + (" ORDER BY " + order_by if order_by else "")
        out = []
        for r in self.db.execute(q, args):
            p = Primer(r["seq"], r["fg_freq"], r["bg_freq"], _py_float(r["ratio"]))
            p._id, p.tm, p.gini, p.active = r["_id"], r["tm"], r["gini"], bool(r["active"])
            p.locations = json.loads(r["_locations"]) if r["_locations"] else None
            out.append(p)
        return out

    def primers_by_seqs(self, seqs):
        seqs = list(seqs)
        out = []
        for i in range(0, len(seqs), 500):
            chunk = seqs[i:i + 500]
            out += self.primers('"seq" IN (%s)' % ", ".join("?" * len(chunk)), chunk)
        return out

    def update_primers(self, primers, fields=("_id", "tm", "_locations", "gini", "active")):
        """Primers._update_in_chunks: write the listed attributes of the records back."""
        def value(p, f):
            if f == "_locations":
                return json.dumps(p.locations) if p.locations is not None else None
            if f == "active":
                return int(bool(p.active))
            return getattr(p, f)
        with self.db:
            self.db.executemany('UPDATE "primer" SET %s WHERE "seq" = ?' % ", ".join('"%s" = ?' % f for f in fields),
                                [[value(p, f) for f in fields] + [p.seq] for p in primers])

    def deactivate_all(self):                       # commands/filter.py:29
        with self.db:
            self.db.execute('UPDATE "primer" SET "active" = 0')

    def reset_ids(self):                            # primers.py:262
        with self.db:
            self.db.execute('UPDATE "primer" SET "_id" = -1')

    # -- sets (Set.add, workspace.py:206-232)
    def count_sets(self):
        return self.db.execute('SELECT COUNT(*) FROM "set"').fetchone()[0]

    def min_set_id(self):                           # commands/score.py:67
        return self.db.execute('SELECT MIN("_id") FROM "set"').fetchone()[0]

    def add_set(self, _id, primer_seqs, score, scoring_fn, **variables):
        """Returns (row id, created).  An equal primer set already in the table is not added again."""
        primer_seqs = list(primer_seqs)
        if not primer_seqs:
            raise ValueError("Cannot add an empty set.")
        if isinstance(score, float) and score != score:
            # `score REAL NOT NULL` (workspace.py:171): a NaN can only come from 0/0 in the statistics, where the
            # reference has already raised (stats.py:53-54 gini with fair_area == 0)
            raise ZeroDivisionError("float division by zero")
        h = set_hash(primer_seqs)
        r = self.db.execute('SELECT "_id" FROM "set" WHERE "_hash" = ?', (h,)).fetchone()
        if r is not None:
            return r[0], False
        cols = ["_id", "_hash", "score", "scoring_fn"] + [v for v in SET_VARIABLES if v in variables]
        vals = [_id, h, _sql_float(score), scoring_fn] + [_sql_float(variables[v]) for v in SET_VARIABLES if v in variables]
        with self.db:
            self.db.execute('INSERT INTO "set" (%s) VALUES (%s)' % (", ".join('"%s"' % c for c in cols),
                                                                   ", ".join("?" * len(cols))), vals)
            self.db.executemany('INSERT INTO "set_primer_through" ("set_id", "primer_id") VALUES (?, ?)',
                                [(_id, s) for s in primer_seqs])
        return _id, True

    def sets(self, order_by='"_id"'):
        out = []
        for r in self.db.execute('SELECT * FROM "set" ORDER BY ' + order_by):
            d = {k: r[k] for k in r.keys()}
            for k in ("score", "bg_dist_mean", "fg_dist_std", "fg_dist_gini", "fg_dist_mean"):
                d[k] = _py_float(d[k])
            d["primers"] = [x[0] for x in self.db.execute(
                'SELECT "primer_id" FROM "set_primer_through" WHERE "set_id" = ? ORDER BY "id"', (r["_id"],))]
            out.append(d)
        return out


def _sql_float(x):
    """inf / nan have no SQLite literal: the reference (peewee -> sqlite3) stores Python floats as they are,
    which sqlite3 binds as REAL inf and NULL for nan."""
    if isinstance(x, float) and x != x:
        return None
    return x


def _py_float(x):
    return float("nan") if x is None else x


def connection(db_name=DEFAULT_DB_FNAME, must_exist=True):
    """workspace.connection (workspace.py:250-262): open, check the version, yield the workspace."""
    if must_exist and not os.path.isfile(db_name):
        raise WorkspaceError("Workspace %s not found: run init first" % db_name)
    ws = Workspace(db_name)
    if must_exist:
        ws.check_version()
    return ws
