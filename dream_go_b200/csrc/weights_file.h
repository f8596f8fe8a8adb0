// Reader of the `dream_go.json` weight file: a JSON object of
//   name -> {"s": base85(f32 max-abs), "t": "i1"|"i4"|"f2"|"f4", "v": base85(little-endian elements)}
// plus plain-string entries that are ignored -- the format of src/libdg_nn/loader.rs:36-116 and
// src/libdg_utils/b85.rs:17-139 (RFC 1924 alphabet, five characters per big-endian 32-bit word).
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

namespace dg {

struct HostTensor {
    std::string dtype;            // "f2", "f4", "i4", "i1"
    float scale = 0.f;            // "s" -- parsed and kept like the reference, never used by forward
    std::vector<uint8_t> bytes;   // decoded elements (may carry base85 padding at the end)
};
typedef std::map<std::string, HostTensor> TensorMap;

// 0 = ok, 1 = missing (unreadable / empty / no tensors), 2 = malformed.
int load_weights_file(const char* path, TensorMap& out, std::string& why);
int parse_weights_json(const char* text, size_t len, TensorMap& out, std::string& why);
bool b85_decode(const char* text, size_t len, std::vector<uint8_t>& out);

}  // namespace dg
