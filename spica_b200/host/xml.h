// Minimal XML DOM for Mitsuba-0.5-style scene files (replaces the vendored tinyxml2 the reference
// uses, spica/sceneparser.cc:72-118): elements, attributes, comments, declarations, self-closing tags.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace spica {
namespace xml {

struct Element {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<Element>> children;
    const char* attribute(const char* key) const {
        for (const auto& a : attrs) if (a.first == key) return a.second.c_str();
        return nullptr;
    }
    bool noChildren() const { return children.empty(); }
};

// Returns the root element, or null with *err set.
std::unique_ptr<Element> parseFile(const std::string& path, std::string* err);
std::unique_ptr<Element> parseString(const std::string& text, std::string* err);

}  // namespace xml
}  // namespace spica
