// schedule.hpp -- host-side planning of one beagleUpdatePartials call.
//
// The reference hands BeagleLib its operation list in reverse level-order
// (Likelihood::defineOperations, reference src/likelihood.hpp:296-355) and BeagleLib executes
// it one operation at a time, re-reading both children from memory.  On B200 that is (n-2)
// dependent passes over HBM.  The planner below regroups the very same operations into
// *chains*: a chain starts at an operation whose children are both already in memory (tips, or
// partials finished in an earlier launch) and then climbs the tree, each next operation taking
// the partials just produced straight from registers plus one sibling from memory.  A chain
// never waits on another chain of the same launch: at a node whose two children sit in chains of
// equal level both chains end and a new one starts one level up (Strahler numbering), so a tree
// of n tips needs at most log2(n) launches instead of one per dependency level, and every
// partials buffer is read from HBM at most once per evaluation.  Every destination buffer is
// still written, so the buffers hold exactly what the reference's would.
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

#include "../../include/strom_beagle.h"

namespace sb200 {

// One operation as the pruning kernel consumes it (48 bytes = three 16-byte words, uploaded as-is).
struct DevOp {
    int dest;    // destination partials slot (= buffer index - tipCount)
    int a;       // operand A: buffer index (tip if < tipCount), or -1 = carried in registers
    int ma;      // transition matrix of A's edge
    int b;       // operand B: buffer index (never carried)
    int mb;      // transition matrix of B's edge
    int sa;      // subtree log-scaler slot to add for A, or -1
    int sb;      // subtree log-scaler slot to add for B, or -1
    int flags;   // OP_* bits
    int sOut;    // OP_CHAIN_END / OP_STORE_S: slot receiving the subtree log-scaler of dest, or -1
    int tt;      // OP_TIPTIP: index of this operation's precomputed (stateA, stateB) table
    int pad[2];
};
enum {
    OP_RESCALE = 1,        // destinationScaleWrite >= 0
    OP_CHAIN_START = 2,    // operand A comes from memory (tips or a finished buffer)
    OP_CHAIN_END = 4,      // last operation of its chain: emit the subtree log-scaler, reset
    OP_ADD_TO_CUM = 8,     // with OP_CHAIN_END: the chain total goes into the cumulative buffer
    OP_TIPTIP = 16,        // both operands are tips: the result is one of 25 precomputed vectors per category
    OP_A_TIP = 32,         // operand A is a tip (only with OP_CHAIN_START)
    OP_B_TIP = 64,         // operand B is a tip
    OP_STORE_S = 128       // plans that keep every node's subtree log-scaler: store the running scaler in slot
                           // sOut and carry on (a chain end stores and resets, as always)
};

// Host-side view of a chain (device code only sees the flags above).
struct DevChain {
    int opStart;   // first DevOp of the chain
    int opCount;
    int sOut;
    int flags;     // CHAIN_ADD_TO_CUM
};
enum { CHAIN_ADD_TO_CUM = 1 };

// One launch.  Its chains are dealt into `groupCount` groups of similar total length; a thread
// block runs all operations of one group, [groupStart[g], groupStart[g+1]), as one pipelined loop.
struct Level {
    int chainStart;
    int chainCount;
    int groupStart;   // index into Schedule::groups
    int groupCount;
    int tipsOnly;     // every operand read from memory is a tip and no operation reads a subtree scaler (level 0 of a
                      // full evaluation): the launch may use the kernel variant without partials-operand code
    int variant;      // thread shape of the pruning kernel this launch runs (chosen by the LevelPolicy; kernels.cuh kVar*)
};

// How a launch is shaped: given the number of independent chains of a level, and whether every operand it reads from
// memory is a tip, the engine picks the kernel's thread shape and how many thread-block groups the chains are dealt into
// (it knows the pattern-tile counts and how many blocks are resident at once).
struct LevelChoice {
    int variant;
    int groups;
};
typedef std::function<LevelChoice(int chainCount, bool tipsCandidate)> LevelPolicy;
// Asked with chainCount == 0 the policy says whether operations on two tips get a precomputed (stateA, stateB) table
// (any answer but groups == kNoTipTipTables), or are computed per pattern from the two tip columns like any other operation
// (small instances: the kernel that builds the tables is 8 us on the critical path of a 90 us evaluation).
constexpr int kNoTipTipTables = -12345;

struct Group {
    int opStart;
    int opEnd;
};

struct Schedule {
    std::vector<DevOp> ops;
    std::vector<DevChain> chains;
    std::vector<Level> levels;
    std::vector<Group> groups;
    std::vector<int> rootSlots;     // slots of forest roots to fold into the cumulative buffer afterwards
    int slotCount = 0;              // subtree log-scaler slots needed
    int tipTipCount = 0;            // operations with OP_TIPTIP (= number of tables)
    bool sequential = false;        // true: list was not a forest; one operation per launch
    // per input operation, for inspection / tests
    std::vector<int> opLevel, opChain;
    // traffic accounting, in units of "one partials buffer" / "one tip row" / "one scaler row"
    long nTipTip = 0, nTipPartial = 0, nPartialPartial = 0;
    long partialReads = 0, partialWrites = 0, tipReads = 0, scalerRows = 0;
};

// Returns BEAGLE_SUCCESS or an error code (index out of range, unsupported scale-read).
// `groupsTipsLevel` (0 = same as groupsPerLevel): the same for a tips-only launch, whose kernel variant has more
// resident blocks per SM.
// `groupsPerLevel`: how many thread-block groups a launch should be split into at most (the engine
// derives it from the pattern-tile count so that every launch has enough blocks to fill the GPU).
// `keepSubtreeScalers`: incremental evaluation.  Slot s = (buffer s + tipCount) permanently holds the total
// log-scaler of the subtree below that buffer: every operation stores its own, and an operand that is an
// internal buffer NOT produced by this list contributes the value an earlier call left in its slot.  The
// list may then cover only the nodes whose partials changed (children before parents, up to the root) and
// the cumulative buffer still receives the whole tree's scaler.  The list must be a forest.
int buildSchedule(const BeagleOperation* ops, int count, int tipCount, int partialsBufferCount,
                  int matrixCount, int scaleCount, bool accumulate, int groupsPerLevel, Schedule& out,
                  bool keepSubtreeScalers = false, int groupsTipsLevel = 0);
// The same with the engine's per-level policy (thread shape and group count from the level's chain count).
int buildSchedule(const BeagleOperation* ops, int count, int tipCount, int partialsBufferCount,
                  int matrixCount, int scaleCount, bool accumulate, const LevelPolicy& policy, Schedule& out,
                  bool keepSubtreeScalers = false);

}  // namespace sb200
