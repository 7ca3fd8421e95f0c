"""Interactive search session: the reference REPL's grammar on top of the GPU path.

Restates the command grammar of image_database.py:2070-2299 (`main()`'s interactive
loop) as a pure parser (``parse_line``) plus a thin loop (``run_session``), so the
``search`` / interactive surface stays drop-in.  Presentation
(HTML gallery, image:/localexplorer: links) is out of scope; results are printed in
the reference's ``"{rank}. {similarity:.4f}: {path}"`` form.

  quit | exit | q                   end the session                          (:2110)
  k:<n>                             number of results                        (:2114-2121)
  folder:<path> | folder:clear      add / clear folder filters               (:2123-2142)
  duplicates:show | duplicates:hide duplicate filter off / on                (:2144-2154)
  <q> - <neg> [- <neg2> ...]        negatives, split on ' - '                (:2157-2190)
  <q1> + <q2>                       combined query, split on the first '+'   (:2193-2213)
  image:<path>                      an image query in any of the positions   (:2167, 2200, 2208, 2227)
  vector:<file.npy>                 (this package) a ready-made float32[1152] embedding in any of the
                                    positions; needs no model: ``vector:q.npy + vector:style.npy - vector:neg.npy``

``python -m clip_database_b200.session --db images.db --embedder mymodule:MyEmbedder`` runs the loop with a
model behind text / image queries (an object with ``text(str)`` and ``image(path)``, see
``database.Embedder``); without ``--embedder`` the session answers ``vector:`` queries only.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple, Union


@dataclass
class SessionState:
    k: int = 10
    weights: Tuple[float, float] = (0.5, 0.5)
    filter_folders: List[str] = field(default_factory=list)
    show_duplicates: bool = False
    profile: bool = False


@dataclass
class Command:
    """A line that changed the session state (or ended it) instead of searching."""
    kind: str          # "quit" | "k" | "folder" | "duplicates" | "empty" | "error"
    message: str = ""


@dataclass
class SearchRequest:
    """Keyword arguments for ``ImageDatabase.search`` (image_database.py:2252-2261)."""
    query: str
    is_image_path: bool = False
    query2: Optional[str] = None
    is_image_path2: bool = False
    negative_query: Optional[str] = None
    negative_is_image: bool = False
    negative_weight: float = 0.5
    negative_queries: Optional[List[str]] = None
    negative_is_images: Optional[List[bool]] = None
    negative_weights: Optional[List[float]] = None

    def kwargs(self, state: SessionState) -> dict:
        return dict(k=state.k, is_image_path=self.is_image_path, query2=self.query2,
                    is_image_path2=self.is_image_path2, weights=state.weights,
                    negative_query=self.negative_query, negative_is_image=self.negative_is_image,
                    negative_weight=self.negative_weight, negative_queries=self.negative_queries,
                    negative_is_images=self.negative_is_images, negative_weights=self.negative_weights,
                    filter_folders=state.filter_folders if state.filter_folders else None,
                    profile=state.profile, show_duplicates=state.show_duplicates)


def _strip_image(part: str) -> Tuple[str, bool]:
    if part.lower().startswith("image:"):
        return part.split(":", 1)[1].strip(), True
    return part, False


def parse_line(line: str, state: SessionState,
               isdir: Callable[[str], bool] = os.path.isdir) -> Union[Command, SearchRequest]:
    """One line of the session.  Mutates ``state`` for the k: / folder: / duplicates:
    commands exactly as the reference loop does; otherwise returns the search to run."""
    query = line.strip()
    if not query:
        return Command("empty")
    low = query.lower()
    if low in ("quit", "exit", "q"):
        return Command("quit", "Ending session. Goodbye!")
    if low.startswith("k:"):
        try:
            state.k = int(query.split(":", 1)[1].strip())
            return Command("k", f"Number of results set to {state.k}")
        except ValueError:
            return Command("error", "Invalid number. Usage: k:20")
    if low.startswith("folder:"):
        folder_path = query.split(":", 1)[1].strip()
        if folder_path.lower() == "clear":
            state.filter_folders = []
            return Command("folder", "Folder filters cleared")
        folder_abs = os.path.abspath(folder_path)
        if not isdir(folder_abs):
            return Command("folder", f"Warning: Folder does not exist: {folder_abs}")
        if folder_abs in state.filter_folders:
            return Command("folder", f"Folder already in filter list: {folder_abs}")
        state.filter_folders.append(folder_abs)
        return Command("folder", f"Added folder filter: {folder_abs}")
    if low.startswith("duplicates:"):
        setting = query.split(":", 1)[1].strip().lower()
        if setting == "show":
            state.show_duplicates = True
            return Command("duplicates", "Duplicate images will be shown")
        if setting == "hide":
            state.show_duplicates = False
            return Command("duplicates", "Duplicate images will be hidden (default)")
        return Command("error", "Invalid option. Use 'duplicates:show' or 'duplicates:hide'")

    req = SearchRequest(query=query)
    # negatives: everything after the first ' - ', further split on ' - '
    if " - " in query:
        head, negative_str = query.split(" - ", 1)
        query = head.strip()
        parts = [p.strip() for p in negative_str.strip().split(" - ")]
        if len(parts) == 1:
            req.negative_query, req.negative_is_image = _strip_image(parts[0])
        else:
            stripped = [_strip_image(p) for p in parts]
            req.negative_queries = [s[0] for s in stripped]
            req.negative_is_images = [s[1] for s in stripped]
            req.negative_weights = [req.negative_weight] * len(stripped)
    # positives: split on the first '+'
    pos = [q.strip() for q in query.split("+", 1)]
    if len(pos) == 2:
        req.query, req.is_image_path = _strip_image(pos[0])
        req.query2, req.is_image_path2 = _strip_image(pos[1])
    else:
        req.query, req.is_image_path = _strip_image(query)
    return req


def run_session(db, state: Optional[SessionState] = None, read: Callable[[str], str] = input,
                write: Callable[[str], None] = print) -> None:
    """The loop of image_database.py:2070-2299 over ``db.search`` (an ``ImageDatabase``)."""
    state = state or SessionState()
    while True:
        try:
            line = read("Query> ")
        except (EOFError, KeyboardInterrupt):
            write("\nEnding session. Goodbye!")
            return
        try:
            item = parse_line(line, state)
            if isinstance(item, Command):
                if item.message:
                    write(item.message)
                if item.kind == "folder" and state.filter_folders:
                    write(f"Current folder filters ({len(state.filter_folders)}):")
                    for f in state.filter_folders:
                        write(f"  - {f}")
                if item.kind == "quit":
                    return
                continue
            results = db.search(item.query, **item.kwargs(state))
            if results:
                write(f"\nFound {len(results)} results:")
                for i, (file_path, similarity) in enumerate(results, 1):
                    write(f"  {i:2d}. {similarity:.4f}: {file_path}")
            else:
                write("No results found.")
            write("")
        except Exception as e:      # the reference keeps the session alive on any error (:2297-2299)
            write(f"Error: {e}")


def main(argv: Optional[List[str]] = None, read: Callable[[str], str] = input,
         write: Callable[[str], None] = print) -> int:
    """``python -m clip_database_b200.session --db images.db [--embedder module:Class]``: search an existing
    database interactively (the reference's ``interactive`` command, image_database.py:2035-2047)."""
    import argparse

    from .database import ImageDatabase, load_embedder
    ap = argparse.ArgumentParser(description="Interactive KNN search over a CLIP-database SQLite file (GPU path)")
    ap.add_argument("--db", required=True)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--devices", default=None, help="comma-separated GPUs to row-shard the store over")
    ap.add_argument("-k", type=int, default=10)
    ap.add_argument("--show-duplicates", action="store_true")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--batch-store", action="store_true", help="keep the bf16 copy: tensor-core pre-selection")
    ap.add_argument("--embedder", default=None,
                    help="module:Class of an embedder (text(str) / image(path) -> float32[1152]); without it only "
                         "vector:<file.npy> queries can be answered")
    args = ap.parse_args(argv)
    embedder = load_embedder(args.embedder) if args.embedder else None
    devices = [int(d) for d in args.devices.split(",")] if args.devices else None
    db = ImageDatabase(args.db, device=args.device, embedder=embedder, verbose=True, batch_store=args.batch_store,
                       devices=devices)
    write("Interactive search: text, image:<path>, vector:<file.npy>; 'q1 + q2', 'q - negative'; k:<n>, "
          "folder:<path>, duplicates:show|hide, quit" if embedder is not None else
          "Interactive search (no embedder: vector:<file.npy> queries only); 'v1 + v2', 'v - negative'; k:<n>, "
          "folder:<path>, duplicates:show|hide, quit")
    try:
        run_session(db, SessionState(k=args.k, show_duplicates=args.show_duplicates, profile=args.profile),
                    read=read, write=write)
    finally:
        db.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
