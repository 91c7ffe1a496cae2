#include "xml.h"

#include <cctype>
#include <fstream>
#include <sstream>

namespace spica {
namespace xml {
namespace {

struct Parser {
    const std::string& s;
    size_t i = 0;
    std::string err;
    explicit Parser(const std::string& t) : s(t) {}

    bool startsWith(const char* p) const { return s.compare(i, std::char_traits<char>::length(p), p) == 0; }
    void skipWs() { while (i < s.size() && std::isspace((unsigned char)s[i])) i++; }
    bool skipMisc() {       // whitespace, comments, declarations, doctype, text
        for (;;) {
            skipWs();
            if (startsWith("<!--")) { const size_t e = s.find("-->", i); if (e == std::string::npos) { err = "unterminated comment"; return false; } i = e + 3; }
            else if (startsWith("<?")) { const size_t e = s.find("?>", i); if (e == std::string::npos) { err = "unterminated declaration"; return false; } i = e + 2; }
            else if (startsWith("<!")) { const size_t e = s.find('>', i); if (e == std::string::npos) { err = "unterminated <!"; return false; } i = e + 1; }
            else if (i < s.size() && s[i] != '<') { while (i < s.size() && s[i] != '<') i++; }
            else return true;
        }
    }
    static std::string unescape(const std::string& v) {
        std::string o;
        for (size_t k = 0; k < v.size(); k++) {
            if (v[k] != '&') { o += v[k]; continue; }
            if (v.compare(k, 4, "&lt;") == 0) { o += '<'; k += 3; }
            else if (v.compare(k, 4, "&gt;") == 0) { o += '>'; k += 3; }
            else if (v.compare(k, 5, "&amp;") == 0) { o += '&'; k += 4; }
            else if (v.compare(k, 6, "&quot;") == 0) { o += '"'; k += 5; }
            else if (v.compare(k, 6, "&apos;") == 0) { o += '\''; k += 5; }
            else o += v[k];
        }
        return o;
    }
    std::string name() {
        const size_t b = i;
        while (i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '-' || s[i] == ':' || s[i] == '.')) i++;
        return s.substr(b, i - b);
    }
    std::unique_ptr<Element> element() {
        if (i >= s.size() || s[i] != '<') { err = "expected '<'"; return nullptr; }
        i++;
        auto e = std::make_unique<Element>();
        e->name = name();
        if (e->name.empty()) { err = "empty element name"; return nullptr; }
        for (;;) {
            skipWs();
            if (i >= s.size()) { err = "unexpected end inside <" + e->name; return nullptr; }
            if (s[i] == '/') { if (i + 1 < s.size() && s[i + 1] == '>') { i += 2; return e; } err = "stray '/'"; return nullptr; }
            if (s[i] == '>') { i++; break; }
            const std::string k = name();
            skipWs();
            if (k.empty() || i >= s.size() || s[i] != '=') { err = "bad attribute in <" + e->name; return nullptr; }
            i++; skipWs();
            if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) { err = "attribute value must be quoted in <" + e->name; return nullptr; }
            const char q = s[i++];
            const size_t b = i;
            while (i < s.size() && s[i] != q) i++;
            if (i >= s.size()) { err = "unterminated attribute value"; return nullptr; }
            e->attrs.emplace_back(k, unescape(s.substr(b, i - b)));
            i++;
        }
        for (;;) {
            if (!skipMisc()) return nullptr;
            if (i >= s.size()) { err = "missing </" + e->name + ">"; return nullptr; }
            if (startsWith("</")) {
                i += 2;
                const std::string n = name();
                skipWs();
                if (n != e->name || i >= s.size() || s[i] != '>') { err = "mismatched </" + n + "> for <" + e->name + ">"; return nullptr; }
                i++;
                return e;
            }
            auto c = element();
            if (!c) return nullptr;
            e->children.push_back(std::move(c));
        }
    }
};

}  // namespace

std::unique_ptr<Element> parseString(const std::string& text, std::string* err) {
    Parser p(text);
    if (!p.skipMisc()) { *err = p.err; return nullptr; }
    auto root = p.element();
    if (!root) *err = p.err + " (offset " + std::to_string(p.i) + ")";
    return root;
}

std::unique_ptr<Element> parseFile(const std::string& path, std::string* err) {
    std::ifstream ifs(path, std::ios::binary);
    if (!ifs) { *err = "cannot open " + path; return nullptr; }
    std::stringstream ss;
    ss << ifs.rdbuf();
    return parseString(ss.str(), err);
}

}  // namespace xml
}  // namespace spica
