// stand-in for Inviwo's DataTraits / Document / utildoc (un-vendored): enough for the DataTraits<PhotonData>
// specialisation in ppm/photondata.h to compile; nothing of it is executed
#pragma once
#include <initializer_list>
#include <string>
#include <utility>
namespace inviwo {
template <typename T>
struct DataTraits;
class Document {
public:
    struct PathComponent {
        static PathComponent end() { return PathComponent(); }
    };
    struct DocumentHandle {};
    DocumentHandle handle() { return DocumentHandle(); }
    void append(const char*, const char*, std::initializer_list<std::pair<const char*, const char*>>) {}
    operator std::string() const { return std::string(); }
};
namespace utildoc {
class TableBuilder {
public:
    struct Header {
        Header(const char*) {}
    };
    TableBuilder(Document::DocumentHandle, Document::PathComponent) {}
    template <typename... A>
    void operator()(A&&...) {}
};
}  // namespace utildoc
}  // namespace inviwo
