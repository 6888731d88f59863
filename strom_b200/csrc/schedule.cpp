// schedule.cpp -- see schedule.hpp.  Pure host code (no CUDA), exercised by the CPU test-suite
// through sb200DebugSchedule.
#include "schedule.hpp"

#include <algorithm>
#include <utility>

namespace sb200 {

static DevOp makeOp(int dest, int a, int ma, int b, int mb, int sa, int sb, int flags, int sOut) {
    DevOp d;
    d.dest = dest; d.a = a; d.ma = ma; d.b = b; d.mb = mb; d.sa = sa; d.sb = sb; d.flags = flags; d.sOut = sOut;
    d.tt = -1;
    d.pad[0] = d.pad[1] = 0;
    return d;
}

// Operations whose two operands are tips produce, per category, one of only 5x5 possible vectors
// (and rescaling maxima): they get a table computed once per evaluation instead of per pattern.
static void assignTipTipTables(Schedule& out, int tipCount, bool tables) {
    out.tipTipCount = 0;
    for (DevOp& d : out.ops) {
        if ((d.flags & OP_CHAIN_START) && d.a >= 0 && d.a < tipCount) d.flags |= OP_A_TIP;
        if (d.b < tipCount) d.flags |= OP_B_TIP;
        if (tables && (d.flags & OP_A_TIP) && (d.flags & OP_B_TIP)) {
            d.flags |= OP_TIPTIP;
            d.tt = out.tipTipCount++;
        }
    }
}

int buildSchedule(const BeagleOperation* ops, int count, int tipCount, int partialsBufferCount,
                  int matrixCount, int scaleCount, bool accumulate, int groupsPerLevel, Schedule& out,
                  bool keepSubtreeScalers, int groupsTipsLevel) {
    if (groupsPerLevel < 1) groupsPerLevel = 1;
    const LevelPolicy fixed = [=](int, bool tipsCandidate) {
        return LevelChoice{0, (groupsTipsLevel > 0 && tipsCandidate) ? groupsTipsLevel : groupsPerLevel};
    };
    return buildSchedule(ops, count, tipCount, partialsBufferCount, matrixCount, scaleCount, accumulate, fixed, out, keepSubtreeScalers);
}

int buildSchedule(const BeagleOperation* ops, int count, int tipCount, int partialsBufferCount,
                  int matrixCount, int scaleCount, bool accumulate, const LevelPolicy& policy, Schedule& out,
                  bool keepSubtreeScalers) {
    out = Schedule();
    const int nbuf = tipCount + partialsBufferCount;
    if (count < 0) return BEAGLE_ERROR_OUT_OF_RANGE;

    for (int o = 0; o < count; ++o) {
        const BeagleOperation& op = ops[o];
        if (op.destinationPartials < tipCount || op.destinationPartials >= nbuf) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (op.child1Partials < 0 || op.child1Partials >= nbuf) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (op.child2Partials < 0 || op.child2Partials >= nbuf) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (op.child1TransitionMatrix < 0 || op.child1TransitionMatrix >= matrixCount) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (op.child2TransitionMatrix < 0 || op.child2TransitionMatrix >= matrixCount) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (op.destinationScaleWrite >= scaleCount) return BEAGLE_ERROR_OUT_OF_RANGE;
        // Per-node scale buffers are write-only in the reference's usage (scale-read is always
        // BEAGLE_OP_NONE, src/likelihood.hpp:329) and are not materialised by this engine.
        if (op.destinationScaleRead != BEAGLE_OP_NONE) return BEAGLE_ERROR_NO_IMPLEMENTATION;
    }

    // Which operation produces each buffer, and how often each produced buffer is consumed.
    std::vector<int> producer(nbuf, -1);
    bool forest = true;
    for (int o = 0; o < count; ++o) {
        int d = ops[o].destinationPartials;
        if (producer[d] >= 0) forest = false;          // written twice
        producer[d] = o;
    }
    std::vector<int> in1(count, -1), in2(count, -1), consumers(count, 0);
    for (int o = 0; o < count && forest; ++o) {
        const BeagleOperation& op = ops[o];
        if (op.child1Partials == op.destinationPartials || op.child2Partials == op.destinationPartials) forest = false;
        int c[2] = {op.child1Partials, op.child2Partials};
        for (int s = 0; s < 2; ++s) {
            int pr = c[s] >= tipCount ? producer[c[s]] : -1;
            if (pr > o) forest = false;                // reads a buffer a later operation overwrites
            if (pr >= 0 && pr < o) {
                (s == 0 ? in1 : in2)[o] = pr;
                if (++consumers[pr] > 1) forest = false;
            }
        }
        if (in1[o] >= 0 && in1[o] == in2[o]) forest = false;
    }

    out.opLevel.assign(count, 0);
    out.opChain.assign(count, 0);
    out.sequential = !forest;

    auto classify = [&](const BeagleOperation& op) {
        bool t1 = op.child1Partials < tipCount, t2 = op.child2Partials < tipCount;
        if (t1 && t2) ++out.nTipTip;
        else if (t1 || t2) ++out.nTipPartial;
        else ++out.nPartialPartial;
    };

    if (!forest && keepSubtreeScalers && accumulate) return BEAGLE_ERROR_NO_IMPLEMENTATION;
    if (!forest) {
        // Conservative path: the list order is the only safe order; one operation per launch.
        for (int o = 0; o < count; ++o) {
            const BeagleOperation& op = ops[o];
            classify(op);
            const bool rescale = op.destinationScaleWrite >= 0;
            int fl = (rescale ? OP_RESCALE : 0) | OP_CHAIN_START | OP_CHAIN_END;
            if (accumulate && rescale) fl |= OP_ADD_TO_CUM;
            out.ops.push_back(makeOp(op.destinationPartials - tipCount, op.child1Partials, op.child1TransitionMatrix,
                                     op.child2Partials, op.child2TransitionMatrix, -1, -1, fl, -1));
            out.chains.push_back(DevChain{o, 1, -1, (accumulate && rescale) ? CHAIN_ADD_TO_CUM : 0});
            out.levels.push_back(Level{o, 1, o, 1, 0, policy(1, false).variant});
            out.groups.push_back(Group{o, o + 1});
            out.opLevel[o] = o;
            out.opChain[o] = o;
            out.partialReads += (op.child1Partials >= tipCount) + (op.child2Partials >= tipCount);
            out.tipReads += (op.child1Partials < tipCount) + (op.child2Partials < tipCount);
            out.partialWrites += 1;
            if (accumulate && rescale) out.scalerRows += 2;
        }
        assignTipTipTables(out, tipCount, policy(0, true).groups != kNoTipTipTables);
        return BEAGLE_SUCCESS;
    }

    // Strahler-style levels and chains.  (Flat arrays throughout: this runs for every new topology of an MCMC, where a
    // vector per chain and per group made the planner as expensive as the evaluation it plans.)
    std::vector<int> chainLevel, chainLen, chainLast;     // per chain: level, operation count, last input operation
    std::vector<int> carriedChild(count, -1);        // which input child (0/1) is carried, -1 none
    std::vector<char> readFromMemory(count, 0);      // produced buffer is re-read from HBM by its parent
    chainLevel.reserve(count);
    chainLen.reserve(count);
    chainLast.reserve(count);
    std::vector<char> levelHasPartialOperand;        // per level: some operation reads an internal buffer from memory
    for (int o = 0; o < count; ++o) {
        classify(ops[o]);
        const int i1 = in1[o], i2 = in2[o];
        const int l1 = i1 >= 0 ? out.opLevel[i1] : -1;
        const int l2 = i2 >= 0 ? out.opLevel[i2] : -1;
        int carry = -1;
        if (i1 >= 0 && (i2 < 0 || l1 > l2)) carry = 0;
        else if (i2 >= 0 && (i1 < 0 || l2 > l1)) carry = 1;
        if (carry >= 0) {
            const int c = carry == 0 ? i1 : i2;
            const int other = carry == 0 ? i2 : i1;
            out.opLevel[o] = out.opLevel[c];
            out.opChain[o] = out.opChain[c];
            ++chainLen[out.opChain[o]];
            chainLast[out.opChain[o]] = o;
            carriedChild[o] = carry;
            if (other >= 0) readFromMemory[other] = 1;
        } else {
            const int lvl = (i1 >= 0) ? l1 + 1 : 0;   // both in-list with equal level, or neither in-list
            out.opLevel[o] = lvl;
            out.opChain[o] = static_cast<int>(chainLen.size());
            chainLen.push_back(1);
            chainLast.push_back(o);
            chainLevel.push_back(lvl);
            if (i1 >= 0) readFromMemory[i1] = 1;
            if (i2 >= 0) readFromMemory[i2] = 1;
        }
        const int lv = out.opLevel[o];
        if (lv >= static_cast<int>(levelHasPartialOperand.size())) levelHasPartialOperand.resize(lv + 1, 0);
        if ((carry != 0 && ops[o].child1Partials >= tipCount) || (carry != 1 && ops[o].child2Partials >= tipCount))
            levelHasPartialOperand[lv] = 1;
    }
    // operations of every chain, bottom-up (= list order), back to back
    const int nChains = static_cast<int>(chainLen.size());
    std::vector<int> chainBegin(nChains + 1, 0), chainFlat(count), chainFill(nChains);
    for (int c = 0; c < nChains; ++c) chainBegin[c + 1] = chainBegin[c] + chainLen[c];
    for (int c = 0; c < nChains; ++c) chainFill[c] = chainBegin[c];
    for (int o = 0; o < count; ++o) chainFlat[chainFill[out.opChain[o]]++] = o;

    // Subtree log-scaler slots: one per chain whose last buffer is re-read from memory by its
    // parent, or that is a root of the forest when there are several roots.
    std::vector<int> slotOfChain(nChains, -1);
    std::vector<int> roots;
    for (int o = 0; o < count; ++o) if (consumers[o] == 0) roots.push_back(o);
    const bool keep = keepSubtreeScalers && accumulate;
    if (keep) {
        // permanent slots, one per partials buffer; every chain end (and, below, every other operation) stores
        for (int c = 0; c < nChains; ++c) slotOfChain[c] = ops[chainLast[c]].destinationPartials - tipCount;
        out.slotCount = partialsBufferCount;
        if (roots.size() > 1)
            for (int r : roots) out.rootSlots.push_back(slotOfChain[out.opChain[r]]);
    } else if (accumulate) {
        for (int c = 0; c < nChains; ++c) {
            const int last = chainLast[c];
            const bool isRoot = consumers[last] == 0;
            if (readFromMemory[last] || (isRoot && roots.size() > 1)) slotOfChain[c] = out.slotCount++;
        }
        if (roots.size() > 1)
            for (int r : roots) out.rootSlots.push_back(slotOfChain[out.opChain[r]]);
    }

    // Emit chains level by level.  Inside a level the chains are dealt, longest first, into the
    // currently lightest of `groupCount` groups (LPT; the lowest-numbered one among equals), and each group's
    // operations are laid out contiguously: a thread block streams through one group.
    int maxLevel = -1;
    for (int l : chainLevel) maxLevel = std::max(maxLevel, l);
    std::vector<int> chainNewId(nChains, -1);
    // chains of every level, in chain order
    std::vector<int> levelBegin(maxLevel + 2, 0), byLevel(nChains);
    for (int c = 0; c < nChains; ++c) ++levelBegin[chainLevel[c] + 1];
    for (int l = 0; l <= maxLevel; ++l) levelBegin[l + 1] += levelBegin[l];
    {
        std::vector<int> fill(levelBegin.begin(), levelBegin.end() - 1);
        for (int c = 0; c < nChains; ++c) byLevel[fill[chainLevel[c]]++] = c;
    }
    out.ops.reserve(count);
    out.chains.reserve(nChains);
    std::vector<int> groupOf(nChains), groupBegin, order;
    std::vector<std::pair<size_t, int>> heap;         // (load, group): min-heap
    for (int lvl = 0; lvl <= maxLevel; ++lvl) {
        int* ids = byLevel.data() + levelBegin[lvl];
        const int nIds = levelBegin[lvl + 1] - levelBegin[lvl];
        std::stable_sort(ids, ids + nIds, [&](int x, int y) { return chainLen[x] > chainLen[y]; });
        const LevelChoice choice = policy(nIds, !levelHasPartialOperand[lvl]);
        const int nGroups = std::max(1, std::min<int>(choice.groups, nIds));
        heap.clear();
        for (int g = 0; g < nGroups; ++g) heap.emplace_back(size_t(0), g);
        const auto heavier = [](const std::pair<size_t, int>& x, const std::pair<size_t, int>& y) { return x > y; };
        // (a vector of (0, g) in increasing g already is a min-heap under `heavier`)
        groupBegin.assign(nGroups + 1, 0);
        for (int q = 0; q < nIds; ++q) {
            std::pop_heap(heap.begin(), heap.end(), heavier);
            std::pair<size_t, int>& lightest = heap.back();
            groupOf[q] = lightest.second;
            ++groupBegin[lightest.second + 1];
            lightest.first += chainLen[ids[q]] + 1;     // +1: a chain start costs about one extra step
            std::push_heap(heap.begin(), heap.end(), heavier);
        }
        for (int g = 0; g < nGroups; ++g) groupBegin[g + 1] += groupBegin[g];
        order.resize(nIds);
        {
            std::vector<int> fill(groupBegin.begin(), groupBegin.end() - 1);
            for (int q = 0; q < nIds; ++q) order[fill[groupOf[q]]++] = ids[q];       // stable: dealing order inside a group
        }
        Level L{static_cast<int>(out.chains.size()), nIds, static_cast<int>(out.groups.size()), nGroups, 0, choice.variant};
        for (int g = 0; g < nGroups; ++g) {
            Group G{static_cast<int>(out.ops.size()), 0};
            for (int q = groupBegin[g]; q < groupBegin[g + 1]; ++q) {
                const int c = order[q];
                const int* cops = chainFlat.data() + chainBegin[c];
                const int clen = chainLen[c];
                DevChain dc{static_cast<int>(out.ops.size()), clen, slotOfChain[c], 0};
                const int last = chainLast[c];
                if (accumulate && consumers[last] == 0 && roots.size() == 1) {
                    dc.flags |= CHAIN_ADD_TO_CUM;
                    out.scalerRows += 2;
                }
                if (dc.sOut >= 0) out.scalerRows += 2;
                chainNewId[c] = static_cast<int>(out.chains.size());
                for (int j = 0; j < clen; ++j) {
                    const int o = cops[j];
                    const BeagleOperation& op = ops[o];
                    // scaler that comes with an operand read from memory: its producer's chain slot, or (incremental
                    // plans) the permanent slot of an internal buffer that an earlier call computed
                    auto slotFor = [&](int inOp, int buffer) {
                        if (!accumulate) return -1;
                        if (inOp >= 0) return slotOfChain[out.opChain[inOp]];
                        return (keep && buffer >= tipCount) ? buffer - tipCount : -1;
                    };
                    int fl = op.destinationScaleWrite >= 0 ? OP_RESCALE : 0;
                    if (j == 0) fl |= OP_CHAIN_START;
                    if (j + 1 == clen) {
                        fl |= OP_CHAIN_END;
                        if (dc.flags & CHAIN_ADD_TO_CUM) fl |= OP_ADD_TO_CUM;
                    }
                    int sOut = (j + 1 == clen) ? dc.sOut : -1;
                    const int dest = op.destinationPartials - tipCount;
                    if (keep && j + 1 < clen) {
                        fl |= OP_STORE_S;
                        sOut = dest;
                        out.scalerRows += 2;
                    }
                    DevOp d;
                    if (carriedChild[o] < 0)
                        d = makeOp(dest, op.child1Partials, op.child1TransitionMatrix, op.child2Partials,
                                   op.child2TransitionMatrix, slotFor(in1[o], op.child1Partials),
                                   slotFor(in2[o], op.child2Partials), fl, sOut);
                    else if (carriedChild[o] == 0)
                        d = makeOp(dest, -1, op.child1TransitionMatrix, op.child2Partials, op.child2TransitionMatrix, -1,
                                   slotFor(in2[o], op.child2Partials), fl, sOut);
                    else
                        d = makeOp(dest, -1, op.child2TransitionMatrix, op.child1Partials, op.child1TransitionMatrix, -1,
                                   slotFor(in1[o], op.child1Partials), fl, sOut);
                    if (d.a >= 0) (d.a < tipCount ? out.tipReads : out.partialReads) += 1;
                    (d.b < tipCount ? out.tipReads : out.partialReads) += 1;
                    out.scalerRows += 2 * ((d.sa >= 0) + (d.sb >= 0));
                    out.partialWrites += 1;
                    out.ops.push_back(d);
                }
                out.chains.push_back(dc);
            }
            G.opEnd = static_cast<int>(out.ops.size());
            out.groups.push_back(G);
        }
        out.levels.push_back(L);
    }
    for (int o = 0; o < count; ++o) out.opChain[o] = chainNewId[out.opChain[o]];
    assignTipTipTables(out, tipCount, policy(0, true).groups != kNoTipTipTables);
    for (Level& L : out.levels) {
        bool tips = L.groupCount > 0;
        for (int g = L.groupStart; g < L.groupStart + L.groupCount && tips; ++g)
            for (int j = out.groups[g].opStart; j < out.groups[g].opEnd && tips; ++j) {
                const DevOp& d = out.ops[j];
                tips = (d.flags & OP_B_TIP) && (!(d.flags & OP_CHAIN_START) || (d.flags & OP_A_TIP)) && d.sa < 0 && d.sb < 0;
            }
        L.tipsOnly = tips ? 1 : 0;
    }
    return BEAGLE_SUCCESS;
}

}  // namespace sb200
