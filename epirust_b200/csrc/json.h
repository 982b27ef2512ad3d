// Minimal JSON reader/writer helpers for the simulation-config JSON (serde_json in the reference).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace epi {

struct JsonValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JsonValue> arr;
    std::vector<std::pair<std::string, JsonValue>> obj;  // insertion order kept

    const JsonValue* find(const std::string& key) const {
        if (kind != Object) return nullptr;
        for (auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const JsonValue& at(const std::string& key) const {
        const JsonValue* v = find(key);
        if (!v) throw std::runtime_error("missing field `" + key + "`");
        return *v;
    }
    double as_number(const std::string& what) const {
        if (kind != Number) throw std::runtime_error("field `" + what + "` must be a number");
        return num;
    }
    uint32_t as_u32(const std::string& what) const {
        const double v = as_number(what);
        if (v < 0 || v > 4294967295.0 || v != (double)(uint64_t)v) throw std::runtime_error("field `" + what + "` must be an unsigned 32-bit integer");
        return (uint32_t)v;
    }
    bool as_bool(const std::string& what) const {
        if (kind != Bool) throw std::runtime_error("field `" + what + "` must be a boolean");
        return b;
    }
    const std::string& as_string(const std::string& what) const {
        if (kind != String) throw std::runtime_error("field `" + what + "` must be a string");
        return str;
    }
};

JsonValue json_parse(const std::string& text);  // throws std::runtime_error with position on malformed input
std::string json_read_file(const std::string& path);

}  // namespace epi
