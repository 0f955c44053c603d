/* ref_assoc_main.cpp — runs the REFERENCE's own StaticFusion::loadAssoc (FrontEnd.cpp:183-214, inside oracle/_ref/libsf_ref.so)
 * in its own process: `ref_assoc <dir> <assocFile>` prints the entry count (or -1), then one line per entry
 * "<timestamp %.17g>\t<depth path>\t<colour path>".  A separate process because iostream code inside a library loaded
 * next to numpy's statically linked libstdc++ crashes on the shared locale objects.  TEST INFRASTRUCTURE. */
#include <cstdio>
extern "C" {
void* ref_create(int);
int ref_load_assoc(void*, const char*, const char*);
double ref_assoc_entry(int, char*, char*, int);
}
int main(int argc, char** argv) {
    if (argc != 3) return 2;
    void* h = ref_create(8);
    const int n = ref_load_assoc(h, argv[1], argv[2]);
    std::printf("%d\n", n);
    static char a[8192], b[8192];
    for (int k = 0; k < n; k++) {
        const double t = ref_assoc_entry(k, a, b, (int)sizeof(a));
        std::printf("%.17g\t%s\t%s\n", t, a, b);
    }
    return 0;
}
