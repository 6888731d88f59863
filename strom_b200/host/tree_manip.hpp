// tree_manip.hpp -- builds and edits trees; its numbering and traversal contract defines the
// operation list the engine receives (mirrors the public surface of reference src/tree_manip.hpp):
//   * tips are numbered name-1 (names are 1-based taxon numbers)            (:264-280)
//   * edge lengths are read with std::stof (float rounding), negatives -> 0 (:292-315)
//   * unrooted trees are re-rooted at tip 0                                 (:513-548, :917)
//   * internal nodes are renumbered nleaves.. in reverse preorder           (:390-402)
//   * level order = breadth first from the root's only child                (:432-473)
#pragma once
#include <cassert>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <memory>
#include <queue>
#include <set>
#include <string>
#include <vector>

#include "tree.hpp"
#include "xstrom.hpp"

namespace strom {

class TreeManip {
  public:
    TreeManip() {}
    explicit TreeManip(Tree::SharedPtr t) { setTree(t); }

    void setTree(Tree::SharedPtr t) { assert(t); _tree = t; }
    Tree::SharedPtr getTree() { return _tree; }
    void clear() { _tree.reset(); }

    double calcTreeLength() const {
        double TL = 0.0;
        for (Node* nd : _tree->_preorder) TL += nd->_edge_length;
        return TL;
    }
    // note: no 1e-12 floor here, exactly like the reference (src/tree_manip.hpp:102-108)
    void scaleAllEdgeLengths(double scaler) {
        for (Node* nd : _tree->_preorder) nd->_edge_length *= scaler;
    }

    std::string makeNewick(unsigned precision) const;
    void buildFromNewick(const std::string& newick, bool rooted, bool allow_polytomies);
    void rerootAt(int node_number);
    void nniNodeSwap(Node* a, Node* b);
    Node* randomInternalEdge(double uniform01);

    typedef std::shared_ptr<TreeManip> SharedPtr;

  private:
    void refreshPreorder();
    void refreshLevelorder();
    static void unlinkChild(Node* parent, Node* child);
    static void appendChild(Node* parent, Node* child);
    static std::string stripComments(const std::string& s);

    Tree::SharedPtr _tree;
};

inline std::string TreeManip::stripComments(const std::string& s) {
    std::string out;
    int depth = 0;
    for (char ch : s) {
        if (ch == '[') ++depth;
        else if (ch == ']') { if (depth > 0) --depth; }
        else if (depth == 0) out += ch;
    }
    return out;
}

inline void TreeManip::unlinkChild(Node* parent, Node* child) {
    if (parent->_left_child == child) parent->_left_child = child->_right_sib;
    else {
        Node* c = parent->_left_child;
        while (c->_right_sib != child) c = c->_right_sib;
        c->_right_sib = child->_right_sib;
    }
    child->_right_sib = nullptr;
    child->_parent = nullptr;
}

inline void TreeManip::appendChild(Node* parent, Node* child) {
    child->_parent = parent;
    child->_right_sib = nullptr;
    if (!parent->_left_child) parent->_left_child = child;
    else {
        Node* c = parent->_left_child;
        while (c->_right_sib) c = c->_right_sib;
        c->_right_sib = child;
    }
}

inline void TreeManip::buildFromNewick(const std::string& newick, bool rooted, bool allow_polytomies) {
    clear();
    _tree.reset(new Tree());
    _tree->_is_rooted = rooted;
    const std::string s = stripComments(newick);

    // count leaves: a name directly after '(' or ','
    unsigned nleaves = 0;
    {
        char prev = 0;
        for (size_t i = 0; i < s.size(); ++i) {
            char ch = s[i];
            if (std::isspace(static_cast<unsigned char>(ch))) continue;
            if ((prev == '(' || prev == ',') && ch != '(' && ch != ',' && ch != ')') ++nleaves;
            prev = ch;
        }
    }
    if (nleaves < 3) throw XStrom("Expecting newick tree description to have at least 3 leaves");
    _tree->_nleaves = nleaves;
    const unsigned max_nodes = 2 * nleaves - (rooted ? 0 : 2);
    _tree->_nodes.resize(max_nodes);
    for (Node& nd : _tree->_nodes) nd._number = -1;

    try {
        unsigned next = 0;
        auto fresh = [&]() -> Node* {
            if (next >= max_nodes)
                throw XStrom("Too many nodes specified by tree description (" + std::to_string(max_nodes) + " nodes allocated for " +
                             std::to_string(nleaves) + " leaves)");
            return &_tree->_nodes[next++];
        };
        Node* nd = fresh();
        _tree->_root = nd;
        if (rooted) {
            Node* below = nd;
            nd = fresh();
            appendChild(below, nd);
        }
        std::set<int> used;
        enum Tok { LParen, RParen, Colon, Comma, Name, EdgeLen };
        Tok prev = LParen;
        bool started = false;
        size_t i = 0;
        const size_t n = s.size();
        auto childCount = [](Node* p) { unsigned c = 0; for (Node* x = p->_left_child; x; x = x->_right_sib) ++c; return c; };
        while (i < n) {
            const char ch = s[i];
            if (std::isspace(static_cast<unsigned char>(ch))) { ++i; continue; }
            if (ch == ';') break;
            if (ch == '(') {
                if (started && !(prev == LParen || prev == Comma))
                    throw XStrom("Not expecting left parenthesis at position " + std::to_string(i + 1) + " in tree description");
                Node* c = fresh();
                appendChild(nd, c);
                nd = c;
                prev = LParen;
                started = true;
                ++i;
            } else if (ch == ')') {
                if (!nd->_parent) throw XStrom("Too many right parentheses at position " + std::to_string(i + 1) + " in tree description");
                if (!(prev == RParen || prev == Name || prev == EdgeLen))
                    throw XStrom("Unexpected right parenthesis at position " + std::to_string(i + 1) + " in tree description");
                nd = nd->_parent;
                if (childCount(nd) < 2)
                    throw XStrom("Internal node has only one child at position " + std::to_string(i + 1) + " in tree description");
                prev = RParen;
                ++i;
            } else if (ch == ',') {
                if (!nd->_parent || !(prev == RParen || prev == Name || prev == EdgeLen))
                    throw XStrom("Unexpected comma at position " + std::to_string(i + 1) + " in tree description");
                Node* par = nd->_parent;
                if (!allow_polytomies) {
                    const unsigned limit = (!par->_parent && !rooted) ? 3u : 2u;   // unrooted: trifurcation at the bottom node
                    if (childCount(par) >= limit)
                        throw XStrom("Polytomy found in the following tree description but polytomies prohibited:\n" + newick);
                }
                Node* c = fresh();
                appendChild(par, c);
                nd = c;
                prev = Comma;
                ++i;
            } else if (ch == ':') {
                if (!(prev == RParen || prev == Name))
                    throw XStrom("Unexpected colon at position " + std::to_string(i + 1) + " in tree description");
                prev = Colon;
                ++i;
            } else {
                // a name (quoted or not) or an edge length
                std::string tok;
                size_t j = i;
                if (ch == '\'') {
                    j = s.find('\'', i + 1);
                    if (j == std::string::npos)
                        throw XStrom("Expecting single quote to mark the end of node name at position " + std::to_string(i + 1) + " in tree description");
                    tok = s.substr(i + 1, j - i - 1);
                    ++j;
                } else {
                    while (j < n && !std::isspace(static_cast<unsigned char>(s[j])) && s[j] != '(' && s[j] != ')' && s[j] != ',' &&
                           s[j] != ':' && s[j] != ';')
                        ++j;
                    tok = s.substr(i, j - i);
                }
                if (prev == Colon) {
                    double d = 0.0;
                    try {
                        size_t used_chars = 0;
                        d = std::stof(tok, &used_chars);     // float precision, like the reference
                        if (used_chars != tok.size()) throw std::invalid_argument(tok);
                    } catch (...) {
                        throw XStrom(tok + " is not interpretable as an edge length");
                    }
                    nd->_edge_length = d < 0.0 ? 0.0 : d;
                    prev = EdgeLen;
                } else {
                    if (!(prev == RParen || prev == LParen || prev == Comma))
                        throw XStrom("Unexpected node name (" + tok + ") at position " + std::to_string(i + 1) + " in tree description");
                    nd->_name = tok;
                    if (!nd->_left_child) {
                        int x = 0;
                        try {
                            size_t used_chars = 0;
                            x = std::stoi(tok, &used_chars);
                            if (used_chars != tok.size()) throw std::invalid_argument(tok);
                        } catch (...) {
                            throw XStrom("node name (" + tok + ") not interpretable as a positive integer");
                        }
                        if (x < 1) throw XStrom("node name (" + tok + ") not interpretable as a positive integer");
                        if (!used.insert(x).second) throw XStrom("leaf number " + std::to_string(x) + " used more than once");
                        nd->_number = x - 1;
                    }
                    prev = Name;
                }
                i = j;
            }
        }
        if (nd != _tree->_root && !(rooted && nd->_parent == _tree->_root))
            throw XStrom("Tree description ended before all parentheses were closed");
        _tree->_nodes.resize(next);     // never grows: pointers stay valid
        if (!rooted) rerootAt(0);
        refreshPreorder();
        refreshLevelorder();
    } catch (const XStrom&) {
        clear();
        throw;
    }
}

// Re-roots at a tip: every node on the path from the tip to the old root becomes a child of its
// former child, taking over the edge length of the edge it now hangs from.
inline void TreeManip::rerootAt(int node_number) {
    Node* tip = nullptr;
    for (Node& c : _tree->_nodes)
        if (c._number == node_number) { tip = &c; break; }
    if (!tip) throw XStrom("no node found with number equal to " + std::to_string(node_number));
    if (tip->_left_child) throw XStrom("cannot currently root trees at internal nodes (e.g. node " + std::to_string(node_number) + ")");

    // path tip = v0, v1 = parent(v0), ..., vk = old root
    std::vector<Node*> path;
    for (Node* x = tip; x; x = x->_parent) path.push_back(x);
    // edge (v_i, v_{i+1}) had its length stored on v_i; afterwards it is stored on v_{i+1}
    std::vector<double> len(path.size());
    for (size_t i = 0; i < path.size(); ++i) len[i] = path[i]->_edge_length;
    // detach every path node from its parent (its other children keep their order) ...
    for (size_t i = 0; i + 1 < path.size(); ++i) unlinkChild(path[i + 1], path[i]);
    // ... and hang each old parent below its former child as the rightmost child, which is what the
    // reference's step-by-step rotation produces (src/tree_manip.hpp:513-648)
    for (size_t i = 0; i + 1 < path.size(); ++i) {
        appendChild(path[i], path[i + 1]);
        path[i + 1]->_edge_length = len[i];
    }
    tip->_parent = nullptr;
    tip->_right_sib = nullptr;
    tip->_edge_length = len.back();
    _tree->_root = tip;
}

inline void TreeManip::refreshPreorder() {
    _tree->_preorder.clear();
    if (!_tree->_root) return;
    _tree->_preorder.reserve(_tree->_nodes.size());
    Node* first = _tree->_root->_left_child;
    assert(first && !first->_right_sib);
    std::vector<Node*> stack(1, first);
    while (!stack.empty()) {
        Node* nd = stack.back();
        stack.pop_back();
        _tree->_preorder.push_back(nd);
        std::vector<Node*> kids;
        for (Node* c = nd->_left_child; c; c = c->_right_sib) kids.push_back(c);
        for (size_t i = kids.size(); i-- > 0;) stack.push_back(kids[i]);
    }
    int curr_internal = static_cast<int>(_tree->_nleaves);
    for (size_t i = _tree->_preorder.size(); i-- > 0;) {
        Node* nd = _tree->_preorder[i];
        if (nd->_left_child) nd->_number = curr_internal++;
    }
    if (_tree->_is_rooted) _tree->_root->_number = curr_internal;
}

inline void TreeManip::refreshLevelorder() {
    _tree->_levelorder.clear();
    if (!_tree->_root) return;
    _tree->_levelorder.reserve(_tree->_nodes.size());
    std::queue<Node*> q;
    q.push(_tree->_root->_left_child);
    while (!q.empty()) {
        Node* nd = q.front();
        q.pop();
        _tree->_levelorder.push_back(nd);
        for (Node* c = nd->_left_child; c; c = c->_right_sib) q.push(c);
    }
}

inline std::string TreeManip::makeNewick(unsigned precision) const {
    char buf[64];
    auto fmtLen = [&](double v) { std::snprintf(buf, sizeof(buf), ":%.*f", static_cast<int>(precision), v); return std::string(buf); };
    // recursive description of the subtree at nd (children in sibling order)
    struct Rec {
        const TreeManip* tm;
        const decltype(fmtLen)* f;
        std::string go(Node* nd) const {
            if (!nd->getLeftChild()) return std::to_string(nd->getNumber() + 1) + (*f)(nd->getEdgeLength());
            std::string s = "(";
            bool first = true;
            for (Node* c = nd->getLeftChild(); c; c = c->getRightSib()) {
                if (!first) s += ",";
                s += go(c);
                first = false;
            }
            return s + ")" + (*f)(nd->getEdgeLength());
        }
    } rec{this, &fmtLen};
    Node* top = _tree->_root->_left_child;
    std::string s = "(";
    if (!_tree->_is_rooted) s += std::to_string(_tree->_root->_number + 1) + fmtLen(top->_edge_length) + ",";
    bool first = true;
    for (Node* c = top->_left_child; c; c = c->_right_sib) {
        if (!first) s += ",";
        s += rec.go(c);
        first = false;
    }
    return s + ")";
}

inline Node* TreeManip::randomInternalEdge(double uniform_deviate) {
    assert(uniform_deviate >= 0.0 && uniform_deviate < 1.0);
    // internal nodes other than preorder[0] own the internal edges (reference src/tree_manip.hpp:960-1011)
    const unsigned num_internal_edges = _tree->_nleaves - 2 - (_tree->_is_rooted ? 0 : 1);
    const unsigned index_of_chosen = 1 + static_cast<unsigned>(std::floor(uniform_deviate * num_internal_edges));
    unsigned visited = 0;
    for (Node* nd : _tree->_preorder)
        if (nd->_left_child) {
            if (visited == index_of_chosen) return nd;
            ++visited;
        }
    assert(false);
    return nullptr;
}

// a: child of x; b: sibling of x (both children of y).  Swap a and b.
inline void TreeManip::nniNodeSwap(Node* a, Node* b) {
    Node* x = a->_parent;
    Node* y = b->_parent;
    assert(x && y && y == x->_parent);
    unlinkChild(x, a);
    unlinkChild(y, b);
    // the reference leaves x as y's first child when b was y's left child, and appends otherwise:
    // both cases end with a the rightmost child of y and b the rightmost child of x
    appendChild(y, a);
    appendChild(x, b);
    refreshPreorder();
    refreshLevelorder();
}

}  // namespace strom
