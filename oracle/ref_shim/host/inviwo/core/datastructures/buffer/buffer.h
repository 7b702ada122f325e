// stand-in for Inviwo's Buffer<T> (un-vendored): only the size bookkeeping ppm/photondata.{h,cpp} uses
#pragma once
#include <cstddef>
namespace inviwo {
template <typename T>
class Buffer {
public:
    size_t getSize() const { return size_; }
    void setSize(size_t n) { size_ = n; }
private:
    size_t size_ = 0;
};
}  // namespace inviwo
