"""``dropin.install`` swaps the reference class's search() for the B200 path and nothing else."""
import inspect
import os

import numpy as np
import pytest

from clip_database_b200 import dropin, synth

import golden_cases
from conftest import have_gpu, tol


class StandInReference:
    """Shaped like the reference's ImageDatabase where the drop-in touches it: ``db_path``, the
    two embedding methods, and a ``search`` with the signature of image_database.py:1308-1317."""

    def __init__(self, db_path, vectors):
        self.db_path = db_path
        self.embedding_dim = 1152
        self._vectors = vectors
        self.calls = 0

    def _get_text_embedding(self, text):
        self.calls += 1
        return self._vectors[text]

    def _get_image_embedding(self, path):
        self.calls += 1
        return self._vectors[path]

    def search(self, query, k=10, is_image_path=False, query2=None, is_image_path2=False, weights=(0.5, 0.5),
               negative_query=None, negative_is_image=False, negative_weight=0.5, negative_queries=None,
               negative_is_images=None, negative_weights=None, filter_folders=None, profile=False,
               show_duplicates=False):
        raise AssertionError("the original search() must not run once the drop-in is installed")


def test_signature_matches_the_reference_when_it_is_available():
    """In the authoring container the real reference can be imported (sqlite_vec stubbed as for
    the golden vectors): the patched method has the reference's exact signature and
    ``uninstall`` puts the original back."""
    if not os.path.exists("/root/reference/image_database.py"):
        pytest.skip("reference sources are only present in the authoring container")
    idb = golden_cases.make_golden.import_reference()
    original = idb.ImageDatabase.search
    try:
        dropin.install(idb.ImageDatabase)
        assert idb.ImageDatabase.search is not original
        got = inspect.signature(idb.ImageDatabase.search).parameters
        want = inspect.signature(original).parameters
        assert [(p.name, p.kind, p.default) for p in got.values()] == \
               [(p.name, p.kind, p.default) for p in want.values()]
        assert idb.ImageDatabase.search.__doc__ == original.__doc__
    finally:
        dropin.uninstall(idb.ImageDatabase)
    assert idb.ImageDatabase.search is original


def test_install_is_idempotent_and_reversible():
    original = StandInReference.search
    try:
        dropin.install(StandInReference)
        dropin.install(StandInReference)
        assert StandInReference._b200_original_search is original
    finally:
        dropin.uninstall(StandInReference)
    assert StandInReference.search is original


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["single_k20", "blend_07_03_negative", "folder_filter", "duplicate_filter_default",
                                  "zero_row_gives_empty", "k_negative_unlimited"])
def test_patched_reference_class_reproduces_reference_outputs(golden, name, tmp_path):
    assert have_gpu()
    case = next(c for c in golden["cases"] if c["name"] == name)
    rows, paths, kwargs, vectors, drop_m, drop_i = golden_cases.inputs_for(case)
    db_path = str(tmp_path / (name + ".db"))
    synth.write_reference_db(db_path, rows, paths, drop_mapping_for=drop_m, drop_image_for=drop_i)
    dropin.install(StandInReference, device=0)
    ref = StandInReference(db_path, vectors)
    try:
        results = ref.search("q1", **kwargs)
        again = ref.search("q1", **kwargs)          # second call: refresh() path, same resident store
        assert results == again and ref.calls >= 2
    finally:
        dropin.uninstall(StandInReference)
        if getattr(ref, "_b200", None) is not None:
            ref._b200.close()
    pos = {p: i for i, p in enumerate(paths)}
    got_pos = [pos[p] for p, _ in results]
    exp_sim = np.array(case["expected_similarities"], dtype=np.float64)
    assert len(got_pos) == len(case["expected_positions"])
    assert np.all(np.abs(np.array([s for _, s in results]) - exp_sim) <= tol(1.0 - exp_sim))
    if name != "blend_07_03_negative":              # blends may swap tolerance ties (checked in test_gpu_golden)
        assert got_pos == case["expected_positions"]
