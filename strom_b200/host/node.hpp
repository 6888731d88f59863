// node.hpp -- left-child / right-sibling tree node (mirrors reference src/node.hpp:16-80).
#pragma once
#include <string>
#include <vector>

namespace strom {

class Tree;
class TreeManip;
class Likelihood;
class Updater;

class Node {
    friend class Tree;
    friend class TreeManip;
    friend class Likelihood;
    friend class Updater;
    friend class TreeUpdater;

  public:
    Node() { clear(); }

    Node* getParent() { return _parent; }
    Node* getLeftChild() { return _left_child; }
    Node* getRightSib() { return _right_sib; }
    int getNumber() const { return _number; }
    std::string getName() const { return _name; }
    double getEdgeLength() const { return _edge_length; }
    // edge lengths set through the node are floored (reference src/node.hpp:76-79)
    void setEdgeLength(double v) { _edge_length = v < smallestEdgeLength() ? smallestEdgeLength() : v; }

    static double smallestEdgeLength() { return 1.0e-12; }   // reference src/main.cpp:11

    typedef std::vector<Node> Vector;
    typedef std::vector<Node*> PtrVector;

  private:
    void clear() {
        _left_child = _right_sib = _parent = nullptr;
        _number = 0;
        _name.clear();
        _edge_length = smallestEdgeLength();
    }

    Node* _left_child;
    Node* _right_sib;
    Node* _parent;
    int _number;
    std::string _name;
    double _edge_length;
};

}  // namespace strom
