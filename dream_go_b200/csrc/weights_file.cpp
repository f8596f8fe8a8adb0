#include "weights_file.h"

#include <cstdio>
#include <cstring>

namespace dg {

static const char kAlphabet[] = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz!#$%&()*+-;<=>?@^_`{|}~";

bool b85_decode(const char* text, size_t len, std::vector<uint8_t>& out) {
    static int8_t table[256];
    static bool init = false;
    if (!init) {
        memset(table, -1, sizeof table);
        for (int i = 0; i < 85; i++) table[static_cast<uint8_t>(kAlphabet[i])] = static_cast<int8_t>(i);
        init = true;
    }
    out.clear();
    out.reserve(len / 5 * 4);
    for (size_t i = 0; i < len; i++)
        if (table[static_cast<uint8_t>(text[i])] < 0) return false;
    for (size_t i = 0; i + 5 <= len; i += 5) {     // a trailing partial group is dropped (b85.rs:101-139)
        uint32_t word = 0;
        for (int j = 0; j < 5; j++) word = word * 85u + static_cast<uint32_t>(table[static_cast<uint8_t>(text[i + j])]);
        out.push_back(static_cast<uint8_t>(word >> 24));
        out.push_back(static_cast<uint8_t>(word >> 16));
        out.push_back(static_cast<uint8_t>(word >> 8));
        out.push_back(static_cast<uint8_t>(word));
    }
    return true;
}

namespace {

struct Cursor {
    const char* p;
    const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool eat(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
    // Returns the raw span of a JSON string (escapes are kept verbatim; tensor names and base85
    // payloads never contain '"' or '\\').
    bool str(const char** s, size_t* n) {
        ws();
        if (p >= end || *p != '"') return false;
        const char* b = ++p;
        while (p < end && *p != '"') { if (*p == '\\') p++; p++; }
        if (p >= end) return false;
        *s = b; *n = static_cast<size_t>(p - b);
        p++;
        return true;
    }
};

}  // namespace

int parse_weights_json(const char* text, size_t len, TensorMap& out, std::string& why) {
    Cursor c{text, text + len};
    out.clear();
    c.ws();
    if (c.p >= c.end) { why = "empty file"; return 1; }
    if (!c.eat('{')) { why = "expected '{'"; return 2; }
    if (!c.eat('}')) {
        do {
            const char* s; size_t n;
            if (!c.str(&s, &n) || !c.eat(':')) { why = "expected \"name\":"; return 2; }
            const std::string name(s, n);
            c.ws();
            if (c.p < c.end && *c.p == '"') {          // plain string entry (e.g. model_name:0) is ignored
                if (!c.str(&s, &n)) { why = "unterminated string"; return 2; }
                continue;
            }
            if (!c.eat('{')) { why = "tensor '" + name + "' is not an object"; return 2; }
            HostTensor t;
            const char* v = nullptr; size_t vn = 0;
            bool have_t = false;
            if (!c.eat('}')) {
                do {
                    const char *k, *val; size_t kn, valn;
                    if (!c.str(&k, &kn) || !c.eat(':') || !c.str(&val, &valn)) { why = "tensor '" + name + "' has a non-string attribute"; return 2; }
                    const std::string key(k, kn);
                    if (key == "s") {
                        std::vector<uint8_t> b;
                        if (!b85_decode(val, valn, b) || b.size() < 4) { why = "tensor '" + name + "': bad scale"; return 2; }
                        memcpy(&t.scale, b.data(), 4);
                    } else if (key == "t") {
                        t.dtype.assign(val, valn);
                        if (t.dtype != "i1" && t.dtype != "i4" && t.dtype != "f2" && t.dtype != "f4") { why = "tensor '" + name + "': unknown type '" + t.dtype + "'"; return 2; }
                        have_t = true;
                    } else if (key == "v") {
                        v = val; vn = valn;
                    } else { why = "tensor '" + name + "': unknown attribute '" + key + "'"; return 2; }
                } while (c.eat(','));
                if (!c.eat('}')) { why = "tensor '" + name + "': expected '}'"; return 2; }
            }
            if (v) {
                if (!have_t) { why = "tensor '" + name + "': value without type"; return 2; }
                if (!b85_decode(v, vn, t.bytes)) { why = "tensor '" + name + "': invalid base85"; return 2; }
            }
            out[name] = std::move(t);
        } while (c.eat(','));
        if (!c.eat('}')) { why = "expected '}' at end of file"; return 2; }
    }
    if (out.empty()) { why = "no tensors in file"; return 1; }     // "an empty result-set is an error" (loader.rs:93-98)
    return 0;
}

int load_weights_file(const char* path, TensorMap& out, std::string& why) {
    FILE* f = fopen(path, "rb");
    if (!f) { why = "cannot open"; return 1; }
    std::vector<char> buf;
    char chunk[1 << 16];
    size_t n;
    while ((n = fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
    fclose(f);
    return parse_weights_json(buf.data(), buf.size(), out, why);
}

}  // namespace dg
