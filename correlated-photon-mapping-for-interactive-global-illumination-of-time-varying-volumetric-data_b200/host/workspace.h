// workspace.h -- reader of Inviwo ".inv" workspace files, host only.
//
// A workspace is the XML serialisation of a processor network: <Processors> (type = class identifier, identifier =
// display name, <Properties> with one <Property identifier=...> per property that differs from its default, nested
// <Properties> for composite properties) and <Connections> between ports (ports carry id="refN", connections refer to
// them with reference="refN").  The reference ships workspaces/CorrelatedPhotonMappingSingleVolume.inv; Inviwo's own
// deserialiser is outside the reference tree, so this is a restatement from that file's structure:
//   scalar        <value content="0.5" />                     (ws:472-474)
//   vector        <value x="1024" y="1024" />                 (ws:444-446)
//   option        <selectedIdentifier content="1/2" />        (ws:555-557)
//   transfer fn   <transferFunction><dataPoints><point><pos x= y= /><rgba x= y= z= w= /></point>...  (ws:498-527)
// The drop-in processors (processors.h) keep the reference's property identifiers, so a workspace written by the real
// application configures them headless.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "inviwo_shim.h"

namespace inviwo {

struct XmlNode {
    std::string tag;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XmlNode>> children;
    const std::string* attr(const std::string& name) const;
    std::string attrOr(const std::string& name, const std::string& dflt) const;
    const XmlNode* child(const std::string& tag) const;
    std::vector<const XmlNode*> childrenNamed(const std::string& tag) const;
};

// minimal XML reader: elements, attributes, comments, declarations; text content is ignored (workspaces carry all
// values in attributes).  Throws std::invalid_argument with a byte offset on malformed input.
std::unique_ptr<XmlNode> parseXml(const std::string& text);

struct WorkspaceProcessor {
    std::string type, identifier;
    const XmlNode* node = nullptr;
    const XmlNode* property(const std::string& path) const;   // "material.anisotropy": nested identifiers
    std::vector<std::string> storedPropertyPaths() const;     // every <Property> that carries a stored value
};

struct WorkspaceConnection {
    int outProcessor = -1, inProcessor = -1;   // indices into Workspace::processors (-1: port not found)
    std::string outPort, inPort;
};

class Workspace {
public:
    static Workspace load(const std::string& path);
    static Workspace parse(const std::string& xml);
    std::vector<WorkspaceProcessor> processors;
    std::vector<WorkspaceConnection> connections;
    std::vector<const WorkspaceProcessor*> ofType(const std::string& type) const;
    // is some outport of a processor of type `outType` connected to inport `inPort` of a processor of type `inType`?
    bool connected(const std::string& outType, const std::string& inType, const std::string& inPort) const;
    std::string describe() const;   // one line per processor and per connection (tests, diagnostics)
private:
    std::shared_ptr<XmlNode> root_;
};

// value readers (false = the node stores no value of that shape: the property keeps its default)
bool readScalar(const XmlNode* prop, double& v);
bool readVec(const XmlNode* prop, double v[4], int& n);
bool readSelected(const XmlNode* prop, std::string& id);
bool readTransferFunction(const XmlNode* prop, std::vector<std::pair<double, vec4>>& points);

// Applies every stored property of `wp` that `proc` also has (by identifier) and returns how many were applied.
// Supported: Float/Int/Bool/IntVec2/IntVec3/IntMinMax, OptionProperty<int>, TransferFunctionProperty,
// AdvancedMaterialProperty (phaseFunction, IOR, roughness, anisotropy), CameraProperty (lookFrom/lookTo/lookUp).
// Unknown identifiers are skipped (GUI-only state); out-of-range values throw std::invalid_argument like Inviwo's
// property validation clamps -- the headless loader refuses instead of guessing.
int applyWorkspaceProperties(Processor& proc, const WorkspaceProcessor& wp, std::vector<std::string>* applied = nullptr);

}  // namespace inviwo
