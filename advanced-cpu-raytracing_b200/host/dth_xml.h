// Minimal XML reader for the reference's scene schema.  The reference uses
// tinyxml2 (src/parser.cpp:28-38); only the subset its parser touches is needed: elements, attributes,
// raw (whitespace-preserved) text of the first text node, comments, declarations.
#pragma once
#include <string>
#include <vector>
#include <memory>

namespace dth {

struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::string text;        // first text node, raw (tinyxml2 GetText() semantics with PRESERVE_WHITESPACE)
    bool has_text = false;
    std::vector<std::unique_ptr<XmlNode>> children;

    const XmlNode* child(const char* n) const {
        for (auto& c : children) if (c->name == n) return c.get();
        return nullptr;
    }
    std::vector<const XmlNode*> children_named(const char* n) const {
        std::vector<const XmlNode*> r;
        for (auto& c : children) if (c->name == n) r.push_back(c.get());
        return r;
    }
    const char* attr(const char* n) const {
        for (auto& a : attrs) if (a.first == n) return a.second.c_str();
        return nullptr;
    }
    bool attr_is(const char* n, const char* v) const {
        const char* a = attr(n);
        return a && std::string(a) == v;
    }
};

// Returns the root element or nullptr (err filled).
std::unique_ptr<XmlNode> xml_parse(const std::string& src, std::string& err);

}  // namespace dth
