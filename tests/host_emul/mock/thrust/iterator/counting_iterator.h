// see cub/device/device_radix_sort.cuh
#pragma once
namespace thrust {
template <class T>
struct counting_iterator {
    T base;
    explicit counting_iterator(T b = T()) : base(b) {}
    T operator[](long long i) const { return (T)(base + (T)i); }
};
}  // namespace thrust
