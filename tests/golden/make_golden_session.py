#!/usr/bin/env python3
"""Generate tests/golden/reference_session.json: the reference's own interactive loop (``main()``'s ``search
--interactive`` session, image_database.py:2070-2299) fed a script of lines, with ``ImageDatabase.search`` replaced by a
recorder.  For every line: what the loop printed as its first message (state commands) or the exact keyword arguments
it called ``search()`` with.  Run in the authoring container only (needs /root/reference).

Pins the session grammar — the ' - ' negatives, the first-'+' split, ``image:`` in every position, ``k:``,
``folder:``, ``duplicates:`` and how their state carries over — with the reference's own parser; tests compare
``clip_database_b200.session.parse_line`` against it."""
from __future__ import annotations

import builtins
import contextlib
import io
import json
import os
import sqlite3
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden  # noqa: E402

OUT = os.path.join(HERE, "reference_session.json")

# {DIR} = an existing directory, {DIR2} = another one (created by the generator and by the test)
SCRIPT = [
    "a red car",
    "  padded query  ",
    "image: /tmp/x.jpg",
    "IMAGE:/tmp/x.jpg",
    "image:/a.jpg + sunset over water",
    "sunset + image:/b.png",
    "image:/a.jpg + image:/b.jpg",
    "colourful design - grey monochrome",
    "design - grey - image:/n.png - abstract",
    "a + b + c - image:/neg.jpg",
    "black-and-white photo",
    "minus -without spaces",
    "x - image:/n1.png",
    "image:/p.jpg - y - z",
    "vector:/tmp/q.npy + vector:/tmp/s.npy - vector:/tmp/n.npy",
    "k: 25",
    "cats",
    "k:abc",
    "K:7",
    "cats again",
    "folder:{DIR}",
    "folder:{DIR}",
    "folder:/definitely/not/here",
    "dogs",
    "folder:{DIR2}",
    "dogs in two folders",
    "folder:clear",
    "dogs anywhere",
    "duplicates:show",
    "birds",
    "duplicates:hide",
    "duplicates:maybe",
    "birds again",
    "",
    "a - b + c",
    "quit",
    "never reached",
]


def main():
    idb = make_golden.import_reference()
    tmp = tempfile.mkdtemp(prefix="golden_session_")
    d1, d2 = os.path.join(tmp, "photos"), os.path.join(tmp, "scans")
    os.mkdir(d1)
    os.mkdir(d2)
    db_path = os.path.join(tmp, "s.db")
    conn = sqlite3.connect(db_path)
    conn.execute("CREATE TABLE images (id INTEGER PRIMARY KEY, file_path TEXT)")
    conn.commit()
    conn.close()
    lines = [ln.replace("{DIR2}", d2).replace("{DIR}", d1) for ln in SCRIPT]
    records = []
    current = {"line": None}

    def fake_init(self, db_path, model_cache_dir=None, *a, **kw):
        self.db_path = db_path

    def fake_search(self, query, **kwargs):
        import copy
        kwargs = copy.deepcopy(dict(kwargs))         # the loop keeps mutating its filter list
        if kwargs.get("weights") is not None:
            kwargs["weights"] = list(kwargs["weights"])
        records[-1]["search"] = {"query": query, **kwargs}
        return []

    feed = iter(lines)

    def fake_input(prompt=""):
        try:
            line = next(feed)
        except StopIteration:
            raise EOFError
        records.append({"line": line, "search": None, "printed": []})
        current["line"] = line
        return line

    class Tee(io.StringIO):
        def write(self, s):
            if records and s.strip():
                records[-1]["printed"].append(s.strip())
            return super().write(s)

    real = (idb.ImageDatabase.__init__, idb.ImageDatabase.search, builtins.input, sys.argv, sys.stdin)
    idb.ImageDatabase.__init__ = fake_init
    idb.ImageDatabase.search = fake_search
    builtins.input = fake_input
    sys.argv = ["image_database.py", "search", "--db", db_path, "--interactive"]

    class Tty(io.StringIO):
        def isatty(self):
            return True
    sys.stdin = Tty()
    try:
        with contextlib.redirect_stdout(Tee()):
            idb.main()
    finally:
        idb.ImageDatabase.__init__, idb.ImageDatabase.search, builtins.input, sys.argv, sys.stdin = real
    out = {"generated_by": "reference image_database.py main() 'search --interactive' loop, search() replaced by a recorder",
           "dir_placeholders": {"{DIR}": "an existing directory", "{DIR2}": "another existing directory"},
           "records": []}
    for rec in records:
        line = rec["line"].replace(d2, "{DIR2}").replace(d1, "{DIR}")
        search = rec["search"]
        if search is not None and search.get("filter_folders"):
            search["filter_folders"] = [f.replace(d2, "{DIR2}").replace(d1, "{DIR}") for f in search["filter_folders"]]
        first = rec["printed"][0].replace(d2, "{DIR2}").replace(d1, "{DIR}") if rec["printed"] else ""
        out["records"].append({"line": line, "search": search, "first_message": None if search is not None else first})
        print(repr(line), "->", search if search is not None else first)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
