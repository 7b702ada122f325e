// workspace.cpp -- see workspace.h
#include "workspace.h"

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "processors.h"

namespace inviwo {

// ---------------------------------------------------------------------------------------------- XML ---------
const std::string* XmlNode::attr(const std::string& name) const {
    for (auto& kv : attrs)
        if (kv.first == name) return &kv.second;
    return nullptr;
}
std::string XmlNode::attrOr(const std::string& name, const std::string& dflt) const {
    const std::string* a = attr(name);
    return a ? *a : dflt;
}
const XmlNode* XmlNode::child(const std::string& t) const {
    for (auto& c : children)
        if (c->tag == t) return c.get();
    return nullptr;
}
std::vector<const XmlNode*> XmlNode::childrenNamed(const std::string& t) const {
    std::vector<const XmlNode*> v;
    for (auto& c : children)
        if (c->tag == t) v.push_back(c.get());
    return v;
}

namespace {

struct XmlReader {
    const std::string& s;
    size_t p = 0;
    explicit XmlReader(const std::string& text) : s(text) {}
    [[noreturn]] void fail(const std::string& what) const {
        throw std::invalid_argument("workspace XML: " + what + " at byte " + std::to_string(p));
    }
    bool startsWith(const char* lit) const { return s.compare(p, std::strlen(lit), lit) == 0; }
    void skipSpace() {
        while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\r' || s[p] == '\n')) ++p;
    }
    void skipUntil(const char* lit) {
        size_t q = s.find(lit, p);
        if (q == std::string::npos) fail(std::string("unterminated construct, expected ") + lit);
        p = q + std::strlen(lit);
    }
    static bool nameChar(char c) {
        return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-' || c == '.' || c == ':';
    }
    std::string name() {
        size_t b = p;
        while (p < s.size() && nameChar(s[p])) ++p;
        if (p == b) fail("expected a name");
        return s.substr(b, p - b);
    }
    static std::string unescape(const std::string& v) {
        if (v.find('&') == std::string::npos) return v;
        static const std::pair<const char*, char> ents[] = {{"&amp;", '&'}, {"&lt;", '<'}, {"&gt;", '>'}, {"&quot;", '"'}, {"&apos;", '\''}};
        std::string out;
        for (size_t i = 0; i < v.size();) {
            bool hit = false;
            if (v[i] == '&')
                for (auto& e : ents)
                    if (v.compare(i, std::strlen(e.first), e.first) == 0) {
                        out += e.second;
                        i += std::strlen(e.first);
                        hit = true;
                        break;
                    }
            if (!hit) out += v[i++];
        }
        return out;
    }
    // skips text, comments, declarations; returns false at end of input
    bool nextTag() {
        while (p < s.size()) {
            size_t q = s.find('<', p);
            if (q == std::string::npos) {
                p = s.size();
                return false;
            }
            p = q;
            if (startsWith("<!--")) skipUntil("-->");
            else if (startsWith("<?")) skipUntil("?>");
            else if (startsWith("<![CDATA[")) skipUntil("]]>");
            else if (startsWith("<!")) skipUntil(">");
            else return true;
        }
        return false;
    }
    std::unique_ptr<XmlNode> element() {   // p at '<' of an opening tag
        ++p;
        auto n = std::make_unique<XmlNode>();
        n->tag = name();
        while (true) {
            skipSpace();
            if (p >= s.size()) fail("unterminated tag <" + n->tag);
            if (s[p] == '/') {
                if (p + 1 >= s.size() || s[p + 1] != '>') fail("expected />");
                p += 2;
                return n;
            }
            if (s[p] == '>') {
                ++p;
                break;
            }
            std::string k = name();
            skipSpace();
            if (p >= s.size() || s[p] != '=') fail("expected = after attribute " + k);
            ++p;
            skipSpace();
            if (p >= s.size() || (s[p] != '"' && s[p] != '\'')) fail("expected a quoted value for " + k);
            const char quote = s[p++];
            size_t e = s.find(quote, p);
            if (e == std::string::npos) fail("unterminated value of " + k);
            n->attrs.emplace_back(k, unescape(s.substr(p, e - p)));
            p = e + 1;
        }
        while (true) {   // children until the matching closing tag
            if (!nextTag()) fail("missing </" + n->tag + ">");
            if (s[p + 1] == '/') {
                p += 2;
                std::string closing = name();
                if (closing != n->tag) fail("</" + closing + "> closes <" + n->tag + ">");
                skipSpace();
                if (p >= s.size() || s[p] != '>') fail("expected > after </" + closing);
                ++p;
                return n;
            }
            n->children.push_back(element());
        }
    }
};

double toDouble(const std::string& v, const char* what) {
    char* end = nullptr;
    double d = std::strtod(v.c_str(), &end);
    if (end == v.c_str() || *end != '\0') throw std::invalid_argument(std::string("workspace: '") + v + "' is not a number (" + what + ")");
    return d;
}

const XmlNode* findProperty(const XmlNode* owner, const std::string& id) {
    const XmlNode* props = owner ? owner->child("Properties") : nullptr;
    if (!props) return nullptr;
    for (auto* p : props->childrenNamed("Property"))
        if (p->attrOr("identifier", "") == id) return p;
    return nullptr;
}

bool storesValue(const XmlNode* p) {
    return p->child("value") || p->child("selectedIdentifier") || p->child("transferFunction");
}

void collectPaths(const XmlNode* owner, const std::string& prefix, std::vector<std::string>& out) {
    const XmlNode* props = owner->child("Properties");
    if (!props) return;
    for (auto* p : props->childrenNamed("Property")) {
        const std::string path = prefix + p->attrOr("identifier", "");
        if (storesValue(p)) out.push_back(path);
        collectPaths(p, path + ".", out);
    }
}

void collectPortIds(const XmlNode* proc, const char* group, int index, std::map<std::string, std::pair<int, std::string>>& ids) {
    const XmlNode* g = proc->child(group);
    if (!g) return;
    for (auto& port : g->children)
        if (const std::string* id = port->attr("id")) ids[*id] = {index, port->attrOr("identifier", "")};
}

}  // namespace

std::unique_ptr<XmlNode> parseXml(const std::string& text) {
    XmlReader r(text);
    if (!r.nextTag()) throw std::invalid_argument("workspace XML: no root element");
    auto root = r.element();
    if (r.nextTag()) r.fail("content after the root element");
    return root;
}

// ---------------------------------------------------------------------------------------- workspace ---------
const XmlNode* WorkspaceProcessor::property(const std::string& path) const {
    const XmlNode* cur = node;
    size_t b = 0;
    while (cur) {
        size_t dot = path.find('.', b);
        cur = findProperty(cur, path.substr(b, dot == std::string::npos ? std::string::npos : dot - b));
        if (dot == std::string::npos) break;
        b = dot + 1;
    }
    return cur;
}
std::vector<std::string> WorkspaceProcessor::storedPropertyPaths() const {
    std::vector<std::string> v;
    collectPaths(node, "", v);
    return v;
}

Workspace Workspace::load(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::invalid_argument("cannot open workspace " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse(ss.str());
}

Workspace Workspace::parse(const std::string& xml) {
    Workspace w;
    w.root_ = std::shared_ptr<XmlNode>(parseXml(xml).release());
    if (w.root_->tag != "InviwoTreeData" && w.root_->tag != "InviwoWorkspace" && w.root_->tag != "ProcessorNetwork")
        throw std::invalid_argument("not an Inviwo workspace: root element <" + w.root_->tag + ">");
    const XmlNode* net = w.root_->child("ProcessorNetwork") ? w.root_->child("ProcessorNetwork") : w.root_.get();
    const XmlNode* procs = net->child("Processors");
    if (!procs) throw std::invalid_argument("workspace has no <Processors>");
    std::map<std::string, std::pair<int, std::string>> portIds;   // refN -> (processor, port identifier)
    for (auto* p : procs->childrenNamed("Processor")) {
        WorkspaceProcessor wp;
        wp.type = p->attrOr("type", "");
        wp.identifier = p->attrOr("identifier", "");
        wp.node = p;
        const int index = (int)w.processors.size();
        collectPortIds(p, "InPorts", index, portIds);
        collectPortIds(p, "OutPorts", index, portIds);
        w.processors.push_back(wp);
    }
    if (const XmlNode* conns = net->child("Connections"))
        for (auto* c : conns->childrenNamed("Connection")) {
            WorkspaceConnection wc;
            const XmlNode *o = c->child("OutPort"), *i = c->child("InPort");
            if (!o || !i) continue;
            wc.outPort = o->attrOr("identifier", "");
            wc.inPort = i->attrOr("identifier", "");
            // either end carries the port's id (first mention) or a reference to it
            for (auto* end : {o, i}) {
                const std::string* ref = end->attr("reference");
                if (!ref) ref = end->attr("id");
                auto it = ref ? portIds.find(*ref) : portIds.end();
                if (it != portIds.end()) (end == o ? wc.outProcessor : wc.inProcessor) = it->second.first;
            }
            w.connections.push_back(wc);
        }
    return w;
}

std::vector<const WorkspaceProcessor*> Workspace::ofType(const std::string& type) const {
    std::vector<const WorkspaceProcessor*> v;
    for (auto& p : processors)
        if (p.type == type) v.push_back(&p);
    return v;
}

bool Workspace::connected(const std::string& outType, const std::string& inType, const std::string& inPort) const {
    for (auto& c : connections)
        if (c.outProcessor >= 0 && c.inProcessor >= 0 && processors[c.outProcessor].type == outType &&
            processors[c.inProcessor].type == inType && c.inPort == inPort)
            return true;
    return false;
}

std::string Workspace::describe() const {
    std::ostringstream os;
    for (auto& p : processors) {
        os << "processor|" << p.type << "|" << p.identifier << "|";
        bool first = true;
        for (auto& path : p.storedPropertyPaths()) {
            os << (first ? "" : ",") << path;
            first = false;
        }
        os << "\n";
    }
    for (auto& c : connections) {
        os << "connection|" << (c.outProcessor >= 0 ? processors[c.outProcessor].identifier : "?") << "." << c.outPort << "|"
           << (c.inProcessor >= 0 ? processors[c.inProcessor].identifier : "?") << "." << c.inPort << "\n";
    }
    return os.str();
}

// ------------------------------------------------------------------------------------------- values ---------
bool readScalar(const XmlNode* prop, double& v) {
    const XmlNode* n = prop ? prop->child("value") : nullptr;
    const std::string* c = n ? n->attr("content") : nullptr;
    if (!c) return false;
    v = toDouble(*c, "scalar value");
    return true;
}
bool readVec(const XmlNode* prop, double v[4], int& n) {
    const XmlNode* node = prop ? prop->child("value") : nullptr;
    if (!node) return false;
    static const char* names[4] = {"x", "y", "z", "w"};
    n = 0;
    for (int k = 0; k < 4; ++k) {
        const std::string* c = node->attr(names[k]);
        if (!c) break;
        v[k] = toDouble(*c, "vector component");
        n = k + 1;
    }
    return n > 0;
}
bool readSelected(const XmlNode* prop, std::string& id) {
    const XmlNode* n = prop ? prop->child("selectedIdentifier") : nullptr;
    const std::string* c = n ? n->attr("content") : nullptr;
    if (!c) return false;
    id = *c;
    return true;
}
bool readTransferFunction(const XmlNode* prop, std::vector<std::pair<double, vec4>>& points) {
    const XmlNode* tf = prop ? prop->child("transferFunction") : nullptr;
    const XmlNode* dp = tf ? tf->child("dataPoints") : nullptr;
    if (!dp) return false;
    points.clear();
    for (auto* pt : dp->childrenNamed("point")) {
        const XmlNode *pos = pt->child("pos"), *rgba = pt->child("rgba");
        if (!pos || !rgba) throw std::invalid_argument("workspace: transfer-function point without <pos>/<rgba>");
        vec4 c(0.f);
        c.x = (float)toDouble(rgba->attrOr("x", "0"), "rgba.x");
        c.y = (float)toDouble(rgba->attrOr("y", "0"), "rgba.y");
        c.z = (float)toDouble(rgba->attrOr("z", "0"), "rgba.z");
        c.w = (float)toDouble(rgba->attrOr("w", "0"), "rgba.w");
        points.emplace_back(toDouble(pos->attrOr("x", "0"), "pos.x"), c);
    }
    return true;
}

namespace {

template <typename T>
void checkRange(const std::string& id, T v, T lo, T hi) {
    if (!(v >= lo && v <= hi)) {
        std::ostringstream os;
        os << "workspace: property '" << id << "' = " << v << " is outside [" << lo << ", " << hi << "]";
        throw std::invalid_argument(os.str());
    }
}

bool applyOne(Property* prop, const XmlNode* node, const std::string& path) {
    double s = 0, v[4] = {0, 0, 0, 0};
    int n = 0;
    std::string sel;
    if (auto* f = dynamic_cast<FloatProperty*>(prop)) {
        if (!readScalar(node, s)) return false;
        checkRange<float>(path, (float)s, f->getMinValue(), f->getMaxValue());
        f->set((float)s);
        return true;
    }
    if (auto* i = dynamic_cast<IntProperty*>(prop)) {
        if (!readScalar(node, s)) return false;
        checkRange<int>(path, (int)std::lround(s), i->getMinValue(), i->getMaxValue());
        i->set((int)std::lround(s));
        return true;
    }
    if (auto* b = dynamic_cast<BoolProperty*>(prop)) {
        if (!readScalar(node, s)) return false;
        b->set(s != 0.0);
        return true;
    }
    if (auto* i2 = dynamic_cast<OrdinalProperty<ivec2>*>(prop)) {   // IntVec2Property and IntMinMaxProperty
        if (!readVec(node, v, n) || n < 2) return false;
        i2->set(ivec2{(int)std::lround(v[0]), (int)std::lround(v[1])});
        return true;
    }
    if (auto* i3 = dynamic_cast<OrdinalProperty<ivec3>*>(prop)) {
        if (!readVec(node, v, n) || n < 3) return false;
        i3->set(ivec3{(int)std::lround(v[0]), (int)std::lround(v[1]), (int)std::lround(v[2])});
        return true;
    }
    if (auto* o = dynamic_cast<OptionProperty<int>*>(prop)) {
        if (!readSelected(node, sel)) return false;
        o->setSelectedIdentifier(sel);   // throws std::invalid_argument for an option the processor does not have
        return true;
    }
    if (auto* t = dynamic_cast<TransferFunctionProperty*>(prop)) {
        std::vector<std::pair<double, vec4>> pts;
        if (!readTransferFunction(node, pts)) return false;
        TransferFunction tf(t->get().getTextureSize());
        for (auto& p : pts) tf.add(p.first, p.second);
        t->set(tf);
        return true;
    }
    if (auto* c = dynamic_cast<CameraProperty*>(prop)) {
        Camera cam = c->get();
        bool any = false;
        for (auto& fld : {std::make_pair("lookFrom", &cam.lookFrom), std::make_pair("lookTo", &cam.lookTo), std::make_pair("lookUp", &cam.lookUp)})
            if (readVec(findProperty(node, fld.first), v, n) && n >= 3) {
                *fld.second = vec3((float)v[0], (float)v[1], (float)v[2]);
                any = true;
            }
        if (any) c->set(cam);
        return any;
    }
    if (auto* m = dynamic_cast<AdvancedMaterialProperty*>(prop)) {
        bool any = false;
        if (readSelected(findProperty(node, "phaseFunction"), sel)) {
            m->phaseFunctionProp.setSelectedIdentifier(sel);
            any = true;
        }
        for (auto& fld : {std::make_pair("IOR", &m->indexOfRefractionProp), std::make_pair("roughness", &m->roughnessProp),
                          std::make_pair("anisotropy", &m->anisotropyProp)})
            if (readScalar(findProperty(node, fld.first), s)) {
                checkRange<float>(path + "." + fld.first, (float)s, fld.second->getMinValue(), fld.second->getMaxValue());
                fld.second->set((float)s);
                any = true;
            }
        if (any) m->propertyModified();
        return any;
    }
    return false;
}

}  // namespace

int applyWorkspaceProperties(Processor& proc, const WorkspaceProcessor& wp, std::vector<std::string>* applied) {
    int count = 0;
    const XmlNode* props = wp.node ? wp.node->child("Properties") : nullptr;
    if (!props) return 0;
    for (auto* node : props->childrenNamed("Property")) {
        const std::string id = node->attrOr("identifier", "");
        Property* prop = proc.getPropertyByIdentifier(id);
        if (!prop) continue;
        if (applyOne(prop, node, wp.identifier + "." + id)) {
            ++count;
            if (applied) applied->push_back(wp.identifier + "." + id);
        }
    }
    return count;
}

}  // namespace inviwo
