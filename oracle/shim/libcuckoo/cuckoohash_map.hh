// Test infrastructure only (oracle/): stand-in for libcuckoo's concurrent hash
// map (absent in this image) with just the member functions the reference
// octree code calls; single-threaded use only.
//
// Iteration order matters: Octree::BalanceFaces (octree.cpp:152-206) tests "is this key still a
// leaf?" while it inserts nodes, so its result depends on the order in which the first pass visits
// the map (found in round 2: at 10 M points 17 of 3.6 M leaves differ between two orders; every
// cloud up to a few million points gives identical trees).  libcuckoo's own order (bucket index of
// the hashed key after its growth history) cannot be reproduced without the library, so the pinned
// oracle uses a DEFINED order: std::map, ascending keys (the default here).  -DASR_SHIM_UNORDERED_MAP
// selects std::unordered_map (round 1's stand-in) for the order-dependence experiment.
#pragma once
#include <cstddef>
#include <map>
#include <unordered_map>

namespace libcuckoo {

template <class K, class V>
class cuckoohash_map {
#ifdef ASR_SHIM_UNORDERED_MAP
    typedef std::unordered_map<K, V> Store;
#else
    typedef std::map<K, V> Store;
#endif
    Store store_;

public:
    class locked_table {
        const Store& s_;

    public:
        explicit locked_table(const Store& s) : s_(s) {}
        typename Store::const_iterator begin() const { return s_.begin(); }
        typename Store::const_iterator end() const { return s_.end(); }
        size_t size() const { return s_.size(); }
        size_t count(const K& k) const { return s_.count(k); }
    };

    locked_table lock_table() { return locked_table(store_); }
    locked_table lock_table() const { return locked_table(store_); }

    bool insert(const K& k, const V& v) { return store_.emplace(k, v).second; }
    bool insert_or_assign(const K& k, const V& v) {
        auto r = store_.emplace(k, v);
        if (!r.second) r.first->second = v;
        return r.second;
    }
    bool find(const K& k, V& out) const {
        auto it = store_.find(k);
        if (it == store_.end()) return false;
        out = it->second;
        return true;
    }
    bool contains(const K& k) const { return store_.count(k) != 0; }
    bool update(const K& k, const V& v) {
        auto it = store_.find(k);
        if (it == store_.end()) return false;
        it->second = v;
        return true;
    }
    size_t size() const { return store_.size(); }
};

}  // namespace libcuckoo
