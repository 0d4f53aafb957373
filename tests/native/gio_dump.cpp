// tests/native/gio_dump.cpp -- TEST INFRASTRUCTURE.
// Reads one lightcone step through the product's column reader (cosmo-cutout_b200/host/column_reader.cpp, compiled into
// this program) and dumps every column as a raw file, so that the GenericIO-format back-end can be checked on a box
// without a GPU.  The reader allocates pinned memory through the C ABI; here those two calls are satisfied by malloc.
// usage: gio_dump <header file> <out dir> [posOnly]
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include "column_reader.h"

extern "C" int cc_host_alloc(void** p, size_t n) { *p = malloc(n); return *p ? 0 : 5; }
extern "C" int cc_host_free(void* p) { free(p); return 0; }
extern "C" const char* cc_last_error(void) { return "malloc failed"; }
namespace cchost {
[[noreturn]] void die(const std::string& msg) { printf("%s\n", msg.c_str()); exit(2); }
}  // namespace cchost
using cchost::die;

static void dump(const std::string& dir, const char* name, const void* p, size_t bytes) {
    FILE* f = fopen((dir + "/" + name + ".bin").c_str(), "wb");
    if (!f || (bytes && fwrite(p, 1, bytes, f) != bytes)) die("cannot write dump");
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 3) die("usage: gio_dump <header> <outdir> [posOnly]");
    const bool pos_only = argc > 3;
    cchost::StepColumns sc;
    sc.read(argv[1], pos_only);
    const std::string out = argv[2];
    dump(out, "x", sc.x.data(), sc.n * 4); dump(out, "y", sc.y.data(), sc.n * 4); dump(out, "z", sc.z.data(), sc.n * 4);
    dump(out, "a", sc.a.data(), sc.n * 4); dump(out, "id", sc.id.data(), sc.n * 8);
    if (!pos_only) {
        dump(out, "vx", sc.vx.data(), sc.n * 4); dump(out, "vy", sc.vy.data(), sc.n * 4); dump(out, "vz", sc.vz.data(), sc.n * 4);
        dump(out, "rotation", sc.rotation.data(), sc.n * 4); dump(out, "replication", sc.replication.data(), sc.n * 4);
    }
    printf("n=%zu\n", sc.n);
    return 0;
}
