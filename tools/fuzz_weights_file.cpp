// AddressSanitizer / UBSan fuzz of the dream_go.json reader (csrc/weights_file.cpp; loader.rs:36-116): 2 M random mutations of a valid
// document, each parsed from an exact-size heap copy -- every outcome is 0 (parsed) / 1 (missing) / 2 (malformed), never a crash.
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 -Idream_go_b200/csrc -include cstring tools/fuzz_weights_file.cpp dream_go_b200/csrc/weights_file.cpp -o /tmp/fuzz_weights && /tmp/fuzz_weights
#include <cstdio>
#include <cstdlib>
#include <string>
#include "weights_file.h"
int main() {
    const std::string good = "{\"model_name:0\": \"x\", \"11v_value/linear_2/offset:0\": {\"s\": \"(^d>V\", \"t\": \"f2\", \"v\": \"(^d>V(^d>V\"}, \"a\": {\"t\": \"i4\", \"v\": \"0000000000\"}, \"b\": {}}";
    dg::TensorMap m; std::string why;
    int rc = dg::parse_weights_json(good.data(), good.size(), m, why);
    printf("valid document: rc %d, %zu tensors\n", rc, m.size());
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&] { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    long counts[3] = {0, 0, 0};
    const char alphabet[] = "{}\":,\\ \n0aZ~[]\x00\xff";
    for (int it = 0; it < 2000000; ++it) {
        std::string doc = good;
        int edits = 1 + rnd() % 4;
        for (int e = 0; e < edits; ++e) {
            switch (rnd() % 4) {
                case 0: doc[rnd() % doc.size()] = alphabet[rnd() % (sizeof(alphabet) - 1)]; break;
                case 1: doc.erase(rnd() % doc.size(), 1 + rnd() % 5); break;
                case 2: doc.insert(rnd() % (doc.size() + 1), 1, alphabet[rnd() % (sizeof(alphabet) - 1)]); break;
                default: doc.resize(rnd() % (doc.size() + 1)); break;
            }
            if (doc.empty()) doc = "{";
        }
        // exact-size heap copy so that AddressSanitizer sees any read past the end
        char* buf = (char*)malloc(doc.size());
        memcpy(buf, doc.data(), doc.size());
        rc = dg::parse_weights_json(buf, doc.size(), m, why);
        free(buf);
        if (rc < 0 || rc > 2) { printf("unexpected rc %d\n", rc); return 1; }
        counts[rc]++;
    }
    printf("2,000,000 mutated documents: %ld parsed, %ld missing, %ld malformed -- no crash, no sanitizer report\n", counts[0], counts[1], counts[2]);
}
