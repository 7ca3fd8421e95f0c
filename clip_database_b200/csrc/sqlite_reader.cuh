// sqlite_reader.cuh — the public SQLite C API, resolved at run time (host code only).
//
// The native loader (clipdb_append_sqlite) reads the reference's database the way the reference's own search
// statement walks it — vec0 in rowid order, joined to image_embeddings and images (image_database.py:1564-1571) —
// with SQLite's C library itself, copying every float32 blob straight from SQLite's page buffer into pinned
// staging memory.  libsqlite3.so.0 is the library Python's sqlite3 module links against; there are no SQLite
// headers in the image, so the handful of public prototypes used are declared here and bound with dlsym.  When
// the library cannot be opened the loader reports CLIPDB_ERR_UNSUPPORTED and the Python reader is used.
#pragma once

#include <dlfcn.h>

#include <mutex>
#include <string>
#include <type_traits>

namespace clipdb {

struct SqliteApi {
    void *dl = nullptr;
    int (*open_v2)(const char *, void **, int, const char *) = nullptr;
    int (*close_v2)(void *) = nullptr;
    int (*prepare_v2)(void *, const char *, int, void **, const char **) = nullptr;
    int (*step)(void *) = nullptr;
    int (*reset)(void *) = nullptr;
    int (*finalize)(void *) = nullptr;
    int (*bind_int64)(void *, int, long long) = nullptr;
    const void *(*column_blob)(void *, int) = nullptr;
    int (*column_bytes)(void *, int) = nullptr;
    long long (*column_int64)(void *, int) = nullptr;
    double (*column_double)(void *, int) = nullptr;
    const unsigned char *(*column_text)(void *, int) = nullptr;
    int (*exec)(void *, const char *, int (*)(void *, int, char **, char **), void *, char **) = nullptr;
    const char *(*errmsg)(void *) = nullptr;
    int (*busy_timeout)(void *, int) = nullptr;
    std::string why;   // why it is unavailable
};

constexpr int SQLITE_OK_ = 0, SQLITE_ROW_ = 100, SQLITE_DONE_ = 101;
constexpr int SQLITE_OPEN_READONLY_ = 0x00000001, SQLITE_OPEN_URI_ = 0x00000040;

inline const SqliteApi *sqlite_api() {
    static SqliteApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libsqlite3.so.0", "libsqlite3.so"};
        for (const char *n : names) {
            api.dl = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.dl) break;
        }
        if (!api.dl) {
            api.why = "libsqlite3.so.0 cannot be opened";
            return;
        }
        bool ok = true;
        auto bind = [&](auto &fn, const char *sym) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.dl, sym));
            if (!fn) {
                ok = false;
                api.why = std::string("libsqlite3 lacks ") + sym;
            }
        };
        bind(api.open_v2, "sqlite3_open_v2");
        bind(api.close_v2, "sqlite3_close_v2");
        bind(api.prepare_v2, "sqlite3_prepare_v2");
        bind(api.step, "sqlite3_step");
        bind(api.reset, "sqlite3_reset");
        bind(api.finalize, "sqlite3_finalize");
        bind(api.bind_int64, "sqlite3_bind_int64");
        bind(api.column_blob, "sqlite3_column_blob");
        bind(api.column_bytes, "sqlite3_column_bytes");
        bind(api.column_int64, "sqlite3_column_int64");
        bind(api.column_double, "sqlite3_column_double");
        bind(api.column_text, "sqlite3_column_text");
        bind(api.exec, "sqlite3_exec");
        bind(api.errmsg, "sqlite3_errmsg");
        bind(api.busy_timeout, "sqlite3_busy_timeout");
        if (!ok) {
            dlclose(api.dl);
            api.dl = nullptr;
        }
    });
    return api.dl ? &api : nullptr;
}

}  // namespace clipdb
