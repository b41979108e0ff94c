// RON (Rusty Object Notation) reader for the reference's scene files.
//
// The reference deserialises `SerializedScene` with the `ron` crate 0.10.1 + serde
// (/root/reference/src/gui/scene_serialized.rs:611-646, written back with
// PrettyConfig::escape_strings(false), :22-24).  There is no Rust toolchain in this image, so
// the host side of portal_b200 reads the same files with this small generic reader: it produces
// a value tree; ph_scene.cpp then walks the tree with the reference's schema.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace ph {

struct RonValue;
using RonPtr = std::shared_ptr<RonValue>;

struct RonValue {
    enum Kind { Null, Bool, Int, Float, String, List, Struct, Map, Tagged } kind = Null;
    bool b = false;
    long long i = 0;
    double f = 0.0;
    std::string s;                                     // String value / Tagged variant name
    std::vector<RonPtr> items;                         // List, tuple; Tagged payload when positional
    std::vector<std::pair<std::string, RonPtr>> fields;  // Struct fields in file order; Tagged struct payload
    std::vector<std::pair<RonPtr, RonPtr>> map;        // Map entries
    bool tagged_struct = false;                        // Tagged: payload is `fields` (true) or `items`

    bool is_num() const { return kind == Int || kind == Float; }
    double num() const { return kind == Int ? double(i) : f; }
    // Struct / tagged-struct field lookup (nullptr when absent)
    const RonValue* get(const std::string& name) const;
    // Positional payload / list element (nullptr when absent)
    const RonValue* at(size_t k) const { return k < items.size() ? items[k].get() : nullptr; }
    bool is_tag(const char* name) const { return kind == Tagged && s == name; }
};

// Parses `text`; on failure returns nullptr and sets `err` ("line N: message").
// `Some(x)` is unwrapped to x and `None` becomes a Null value, as serde's Option does.
RonPtr ron_parse(const std::string& text, std::string& err);

}  // namespace ph
