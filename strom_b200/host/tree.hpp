// tree.hpp -- tree container (mirrors reference src/tree.hpp:14-80): unrooted trees are "rooted" at
// a tip whose only child is preorder[0]; Likelihood reads the traversal vectors directly.
#pragma once
#include <atomic>
#include <memory>

#include "node.hpp"

namespace strom {

class Tree {
    friend class TreeManip;
    friend class Likelihood;
    friend class Updater;
    friend class TreeUpdater;

  public:
    Tree() : _uid(nextUid()) { clear(); }
    Tree(const Tree& other) = delete;             // (node pointers refer into _nodes)
    Tree& operator=(const Tree& other) = delete;
    // Process-wide unique identity of this Tree object (never reused, unlike its address): what a Likelihood that keeps
    // device state between calls remembers about "the tree the state was computed from".
    unsigned long long uid() const { return _uid; }
    bool isRooted() const { return _is_rooted; }
    unsigned numLeaves() const { return _nleaves; }
    unsigned numNodes() const { return static_cast<unsigned>(_nodes.size()); }

    typedef std::shared_ptr<Tree> SharedPtr;

  private:
    void clear() {
        _is_rooted = false;
        _root = nullptr;
        _nleaves = 0;
        _nodes.clear();
        _preorder.clear();
        _levelorder.clear();
    }

    static unsigned long long nextUid() {
        static std::atomic<unsigned long long> counter(0);
        return ++counter;
    }

    const unsigned long long _uid;
    bool _is_rooted;
    Node* _root;
    unsigned _nleaves;
    Node::PtrVector _preorder;
    Node::PtrVector _levelorder;
    Node::Vector _nodes;
};

}  // namespace strom
