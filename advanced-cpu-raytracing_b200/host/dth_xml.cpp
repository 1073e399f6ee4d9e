#include "dth_xml.h"
#include <cctype>
#include <cstring>

namespace dth {

namespace {
struct P {
    const std::string& s;
    size_t i = 0;
    std::string err;
    explicit P(const std::string& src) : s(src) {}
    bool eof() const { return i >= s.size(); }
    bool starts(const char* t) const { return s.compare(i, strlen(t), t) == 0; }
    void skip_ws() { while (!eof() && isspace((unsigned char)s[i])) i++; }

    static std::string decode(const std::string& t) {
        if (t.find('&') == std::string::npos) return t;
        std::string o;
        for (size_t k = 0; k < t.size(); k++) {
            if (t[k] == '&') {
                if (t.compare(k, 4, "&lt;") == 0) { o += '<'; k += 3; continue; }
                if (t.compare(k, 4, "&gt;") == 0) { o += '>'; k += 3; continue; }
                if (t.compare(k, 5, "&amp;") == 0) { o += '&'; k += 4; continue; }
                if (t.compare(k, 6, "&quot;") == 0) { o += '"'; k += 5; continue; }
                if (t.compare(k, 6, "&apos;") == 0) { o += '\''; k += 5; continue; }
            }
            o += t[k];
        }
        return o;
    }

    // skip comments, declarations, doctype; returns false on malformed input
    bool skip_misc() {
        for (;;) {
            skip_ws();
            if (starts("<!--")) {
                size_t e = s.find("-->", i + 4);
                if (e == std::string::npos) { err = "unterminated comment"; return false; }
                i = e + 3;
            } else if (starts("<?")) {
                size_t e = s.find("?>", i + 2);
                if (e == std::string::npos) { err = "unterminated declaration"; return false; }
                i = e + 2;
            } else if (starts("<!")) {
                size_t e = s.find('>', i);
                if (e == std::string::npos) { err = "unterminated <!"; return false; }
                i = e + 1;
            } else return true;
        }
    }

    std::unique_ptr<XmlNode> element() {
        if (eof() || s[i] != '<') { err = "expected '<'"; return nullptr; }
        i++;
        std::unique_ptr<XmlNode> n(new XmlNode());
        size_t b = i;
        while (!eof() && !isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/') i++;
        n->name = s.substr(b, i - b);
        // attributes
        for (;;) {
            skip_ws();
            if (eof()) { err = "eof in tag"; return nullptr; }
            if (s[i] == '/') {
                if (i + 1 < s.size() && s[i + 1] == '>') { i += 2; return n; }
                err = "bad '/'"; return nullptr;
            }
            if (s[i] == '>') { i++; break; }
            size_t ab = i;
            while (!eof() && s[i] != '=' && !isspace((unsigned char)s[i])) i++;
            std::string an = s.substr(ab, i - ab);
            skip_ws();
            if (eof() || s[i] != '=') { err = "attribute without '='"; return nullptr; }
            i++;
            skip_ws();
            if (eof() || (s[i] != '"' && s[i] != '\'')) { err = "attribute value not quoted"; return nullptr; }
            char q = s[i++];
            size_t vb = i;
            while (!eof() && s[i] != q) i++;
            if (eof()) { err = "unterminated attribute"; return nullptr; }
            n->attrs.emplace_back(an, decode(s.substr(vb, i - vb)));
            i++;
        }
        // content
        for (;;) {
            if (eof()) { err = "eof in element " + n->name; return nullptr; }
            if (s[i] == '<') {
                if (starts("<!--")) {
                    size_t e = s.find("-->", i + 4);
                    if (e == std::string::npos) { err = "unterminated comment"; return nullptr; }
                    i = e + 3;
                    continue;
                }
                if (starts("</")) {
                    size_t e = s.find('>', i);
                    if (e == std::string::npos) { err = "unterminated close tag"; return nullptr; }
                    i = e + 1;
                    return n;
                }
                if (starts("<![CDATA[")) {
                    size_t e = s.find("]]>", i);
                    if (e == std::string::npos) { err = "unterminated CDATA"; return nullptr; }
                    if (!n->has_text && n->children.empty()) { n->text = s.substr(i + 9, e - i - 9); n->has_text = true; }
                    i = e + 3;
                    continue;
                }
                auto c = element();
                if (!c) return nullptr;
                n->children.push_back(std::move(c));
            } else {
                size_t tb = i;
                while (!eof() && s[i] != '<') i++;
                // tinyxml2: GetText() is the FIRST child if it is a text node; whitespace-only text
                // between elements is discarded.
                std::string t = s.substr(tb, i - tb);
                bool all_ws = true;
                for (char ch : t) if (!isspace((unsigned char)ch)) { all_ws = false; break; }
                if (!all_ws && !n->has_text && n->children.empty()) { n->text = decode(t); n->has_text = true; }
            }
        }
    }
};
}  // namespace

std::unique_ptr<XmlNode> xml_parse(const std::string& src, std::string& err) {
    P p(src);
    if (!p.skip_misc()) { err = p.err; return nullptr; }
    auto r = p.element();
    if (!r) err = p.err;
    return r;
}

}  // namespace dth
