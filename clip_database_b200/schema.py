"""The reference's SQLite schema, kept verbatim as the loader's input contract.

  images            image_database.py:275-283
  vec0              image_database.py:290-294  (virtual table ``USING vec0(embedding float[1152])``)
  image_embeddings  image_database.py:308-314  (rowid mirrors vec0.rowid, :1178-1181)
  binary_embeddings image_database.py:320-331  (1152 sign bytes per image, :1189-1195)

sqlite-vec is not installed in this image, so synthetic databases use a plain
table of the same name and columns as a stand-in for the virtual table
(``VEC0_STANDIN``); the loader also understands sqlite-vec's shadow tables.
"""
EMBEDDING_DIM = 1152  # SigLIP 2 SO400M; image_database.py:290-294

IMAGES = """
CREATE TABLE IF NOT EXISTS images (
    id INTEGER PRIMARY KEY AUTOINCREMENT,
    file_path TEXT UNIQUE NOT NULL,
    last_modified REAL NOT NULL,
    file_hash TEXT,
    created_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP
)"""

VEC0_VIRTUAL = "CREATE VIRTUAL TABLE IF NOT EXISTS vec0 USING vec0(embedding float[{dim}])"

VEC0_STANDIN = "CREATE TABLE IF NOT EXISTS vec0 (rowid INTEGER PRIMARY KEY, embedding BLOB)"

IMAGE_EMBEDDINGS = """
CREATE TABLE IF NOT EXISTS image_embeddings (
    rowid INTEGER PRIMARY KEY,
    image_id INTEGER,
    FOREIGN KEY (image_id) REFERENCES images(id)
)"""

BINARY_EMBEDDINGS = """
CREATE TABLE IF NOT EXISTS binary_embeddings (
    rowid INTEGER PRIMARY KEY AUTOINCREMENT,
    image_id INTEGER UNIQUE NOT NULL,
    embedding BLOB NOT NULL,
    FOREIGN KEY (image_id) REFERENCES images(id)
)"""

BINARY_EMBEDDINGS_INDEX = """
CREATE INDEX IF NOT EXISTS idx_binary_embeddings_image_id
ON binary_embeddings(image_id)"""

# sqlite-vec's own storage for a table named vec0 [UPSTREAM-UNVERIFIED, recalled from
# sqlite-vec 0.1.x]: rows live in fixed-size chunks (default 1024 rows); per chunk a
# validity bitmap, an int64 rowid array and one packed float32 blob per vector column.
SHADOW_CHUNKS = """
CREATE TABLE IF NOT EXISTS vec0_chunks (
    chunk_id INTEGER PRIMARY KEY AUTOINCREMENT,
    size INTEGER NOT NULL,
    validity BLOB NOT NULL,
    rowids BLOB NOT NULL
)"""
SHADOW_ROWIDS = """
CREATE TABLE IF NOT EXISTS vec0_rowids (
    rowid INTEGER PRIMARY KEY AUTOINCREMENT,
    id,
    chunk_id INTEGER,
    chunk_offset INTEGER
)"""
SHADOW_VECTORS = """
CREATE TABLE IF NOT EXISTS vec0_vector_chunks00 (
    rowid PRIMARY KEY,
    vectors BLOB NOT NULL
)"""
