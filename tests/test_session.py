"""Grammar of the interactive session (image_database.py:2110-2239), host logic only."""
from clip_database_b200.session import Command, SearchRequest, SessionState, parse_line, run_session


def test_plain_text_and_image_queries():
    st = SessionState()
    r = parse_line("  a red car ", st)
    assert isinstance(r, SearchRequest) and r.query == "a red car" and not r.is_image_path and r.query2 is None
    r = parse_line("image: /tmp/x.jpg", st)
    assert r.query == "/tmp/x.jpg" and r.is_image_path
    r = parse_line("IMAGE:/tmp/x.jpg", st)
    assert r.query == "/tmp/x.jpg" and r.is_image_path


def test_combined_and_negative_forms():
    st = SessionState()
    r = parse_line("image:/a.jpg + sunset over water", st)
    assert (r.query, r.is_image_path, r.query2, r.is_image_path2) == ("/a.jpg", True, "sunset over water", False)
    r = parse_line("colourful design - grey monochrome", st)
    assert r.query == "colourful design" and r.negative_query == "grey monochrome" and r.negative_queries is None
    r = parse_line("design - grey - image:/n.png - abstract", st)
    assert r.query == "design" and r.negative_query is None
    assert r.negative_queries == ["grey", "/n.png", "abstract"]
    assert r.negative_is_images == [False, True, False] and r.negative_weights == [0.5, 0.5, 0.5]
    r = parse_line("a + b + c - image:/neg.jpg", st)
    assert (r.query, r.query2) == ("a", "b + c") and r.negative_query == "/neg.jpg" and r.negative_is_image
    # a hyphen without surrounding spaces is not a negative
    r = parse_line("black-and-white photo", st)
    assert r.query == "black-and-white photo" and r.negative_query is None


def test_state_commands(tmp_path):
    st = SessionState()
    assert parse_line("k: 25", st).kind == "k" and st.k == 25
    assert parse_line("k:abc", st).kind == "error" and st.k == 25
    d = tmp_path / "photos"
    d.mkdir()
    assert "Added" in parse_line(f"folder:{d}", st).message and st.filter_folders == [str(d)]
    assert "already" in parse_line(f"folder:{d}", st).message and len(st.filter_folders) == 1
    assert "does not exist" in parse_line("folder:/definitely/not/here", st).message
    assert parse_line("folder:clear", st).kind == "folder" and st.filter_folders == []
    parse_line("duplicates:show", st)
    assert st.show_duplicates
    parse_line("Duplicates:HIDE", st)
    assert not st.show_duplicates
    assert parse_line("duplicates:maybe", st).kind == "error"
    assert parse_line("", st).kind == "empty"
    for word in ("quit", "EXIT", "q"):
        assert parse_line(word, st).kind == "quit"


def test_session_loop_calls_search_with_reference_kwargs():
    calls = []

    class FakeDb:
        def search(self, query, **kw):
            calls.append((query, kw))
            return [("/p/a.jpg", 0.91234), ("/p/b.jpg", 0.5)] if query != "nothing" else []

    lines = iter(["k:3", "duplicates:show", "cat + image:/d.jpg - dog", "nothing", "quit"])
    out = []
    run_session(FakeDb(), SessionState(weights=(0.7, 0.3)), read=lambda _: next(lines), write=out.append)
    assert len(calls) == 2
    q, kw = calls[0]
    assert q == "cat" and kw["k"] == 3 and kw["query2"] == "/d.jpg" and kw["is_image_path2"]
    assert kw["negative_query"] == "dog" and kw["weights"] == (0.7, 0.3) and kw["show_duplicates"]
    assert kw["filter_folders"] is None
    assert any("0.9123: /p/a.jpg" in o for o in out) and "No results found." in out


def test_vector_queries_pass_through_the_grammar_unchanged():
    """``vector:<file.npy>`` needs no parser support: it travels as the query text and ImageDatabase resolves it."""
    st = SessionState()
    r = parse_line("vector:/tmp/q.npy + vector:/tmp/s.npy - vector:/tmp/n.npy", st)
    assert (r.query, r.is_image_path, r.query2, r.is_image_path2) == ("vector:/tmp/q.npy", False, "vector:/tmp/s.npy", False)
    assert r.negative_query == "vector:/tmp/n.npy" and not r.negative_is_image


def test_load_vector_and_embedder_hook(tmp_path):
    import numpy as np
    import pytest
    from clip_database_b200.database import load_embedder, load_vector
    v = np.arange(1152, dtype=np.float64)
    np.save(tmp_path / "q.npy", v)
    got = load_vector(str(tmp_path / "q.npy"), 1152)
    assert got.dtype == np.float32 and np.array_equal(got, v.astype(np.float32))
    np.save(tmp_path / "short.npy", v[:10])
    with pytest.raises(ValueError):
        load_vector(str(tmp_path / "short.npy"), 1152)
    emb = load_embedder("fake_embedder:HashEmbedder")
    assert emb.text("a red car").shape == (1152,) and not np.array_equal(emb.text("a"), emb.image("a"))
    with pytest.raises(TypeError):
        load_embedder("fake_embedder:NotAnEmbedder")
    with pytest.raises(ValueError):
        load_embedder("no_colon")


def test_grammar_matches_the_references_own_loop(tmp_path):
    """tests/golden/reference_session.json: the reference's interactive loop (image_database.py:2070-2299) was fed this
    script with ``search()`` replaced by a recorder (tests/golden/make_golden_session.py).  ``parse_line`` must turn
    every line into the same call — or the same message — with the same state carried from line to line."""
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    golden = json.load(open(os.path.join(here, "golden", "reference_session.json")))
    d1, d2 = tmp_path / "photos", tmp_path / "scans"
    d1.mkdir()
    d2.mkdir()

    def real(s):
        return s.replace("{DIR2}", str(d2)).replace("{DIR}", str(d1)) if isinstance(s, str) else s
    st = SessionState()
    searches = 0
    for rec in golden["records"]:
        item = parse_line(real(rec["line"]), st)
        want = rec["search"]
        if want is None:
            assert isinstance(item, Command), rec["line"]
            assert item.message == real(rec["first_message"]), rec["line"]
            if item.kind == "quit":
                break
            continue
        searches += 1
        assert isinstance(item, SearchRequest), rec["line"]
        got = dict(query=item.query, **item.kwargs(st))
        got["weights"] = list(got["weights"])
        want = dict(want)
        want["filter_folders"] = [real(f) for f in want["filter_folders"]] if want["filter_folders"] else None
        assert got == want, rec["line"]
    assert searches == 23 and item.kind == "quit"
