// Host-side, one-time construction of the node-block sparse structure that every kernel
// works on (DESIGN.md "Data layout").  Input is exactly what the reference's setup()
// produces with deal.II (reference src/classes/NavierStokes.cpp:83-104, 256-273): the
// per-cell global DoF indices in FESystem local order after the component-wise
// renumbering, n_u, n_p, and the vertex coordinates.  The sparsity pattern is the union
// of the dense cell couplings (make_sparsity_pattern, keep_constrained_dofs = true),
// stored compressed at node granularity: the velocity DoFs of one P2 node are
// consecutive (dim*k .. dim*k+dim-1), so one neighbour-node index stands for dim columns.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace nsb {

struct Structure {
  int dim = 0, NV = 0, NN = 0, DPC = 0;
  int rank = 0, nranks = 1;
  // global sizes
  int64_t n_u_glob = 0, n_p_glob = 0, n_cells_glob = 0;
  // local sizes: nodes = P2 nodes (velocity), pids = pressure DoFs (vertices)
  int nn_own = 0, nn_ghost = 0, np_own = 0, np_ghost = 0, nc = 0;
  std::vector<int64_t> node_gid;     // [nn_own+nn_ghost] global node id (= global velocity dof / dim)
  std::vector<int64_t> pid_gid;      // [np_own+np_ghost] global pressure id (= global dof - n_u)
  std::vector<int> cell_gid;         // [nc] global cell index
  std::vector<int> cell_nodes;       // [nc][NN] local node ids
  std::vector<int> cell_pids;        // [nc][NV] local pressure ids
  std::vector<int> node_pid;         // [nn_tot] local pressure id of a vertex node, -1 for line nodes
  std::vector<int> pid_node;         // [np_tot] local node id of the vertex carrying this pressure DoF
  // adjacency of owned nodes, neighbours sorted by global id (=> ascending global column)
  std::vector<int64_t> nbr_ptr;      // [nn_own+1]
  std::vector<int> nbr;              // local node ids
  std::vector<int64_t> pnbr_ptr;     // [nn_own+1]
  std::vector<int> pnbr;             // local pressure ids
  std::vector<int> selfrank;         // [nn_own] position of the node in its own nbr list
  std::vector<int> pselfrank;        // [np_own] position of the pressure id in pnbr of its vertex node
  std::vector<int64_t> rowbase;      // [nn_own] offset of scalar row (node,0) in the value array
  std::vector<int64_t> prowbase;     // [np_own] offset of the pressure row
  int64_t nnz_local = 0;
  // node -> cells (ascending local cell), packed cell<<4 | local node index
  std::vector<int64_t> n2c_ptr;      // [nn_own+1]
  std::vector<uint32_t> n2c;
  // per cell: position of node b / vertex j in the neighbour lists of node a
  std::vector<uint16_t> rank_uu;     // [nc][NN][NN]
  std::vector<uint16_t> rank_up;     // [nc][NN][NV]
  // per cell geometry (constant in time): grad lambda_k [NV][DIM], |det J|, diameter h; 16 doubles/cell
  std::vector<double> cell_geom;
  // halo plan (multi-GPU): ghosts are ordered by (owner rank, global id)
  std::vector<int> peer;                       // ranks this rank exchanges with
  std::vector<std::vector<int>> send_nodes;    // per peer: owned local node ids to send
  std::vector<std::vector<int>> send_pids;     // per peer: owned local pressure ids to send
  std::vector<int> recv_node_count, recv_pid_count;   // per peer, contiguous in the ghost ranges
  int max_row_len = 0, max_node_smem_doubles = 0;

  // vector layout: [u_own (dim*nn_own) | p_own | u_ghost | p_ghost]
  int64_t n_own_dofs() const { return (int64_t)dim * nn_own + np_own; }
  int64_t n_tot_dofs() const { return (int64_t)dim * (nn_own + nn_ghost) + np_own + np_ghost; }
  int64_t node_xoff(int A) const {
    return A < nn_own ? (int64_t)dim * A : n_own_dofs() + (int64_t)dim * (A - nn_own);
  }
  int64_t pid_xoff(int P) const {
    return P < np_own ? (int64_t)dim * nn_own + P
                      : n_own_dofs() + (int64_t)dim * nn_ghost + (P - np_own);
  }
  int row_len(int A) const {
    return dim * (int)(nbr_ptr[A + 1] - nbr_ptr[A]) + (int)(pnbr_ptr[A + 1] - pnbr_ptr[A]);
  }
};

// Builds the structure for `rank` of `nranks`.  cell_part may be null (single rank).
// Returns an empty string on success, else an error message.
std::string build_structure(int dim, int64_t n_vertices, const double* coords, int64_t n_cells,
                            const uint32_t* cell_vertices, const uint32_t* cell_dofs, int64_t n_u,
                            int64_t n_p, const int32_t* cell_part, int rank, int nranks, Structure& S);

// Plan of the SpMV tiles (linalg.cuh / velstream.cuh): runs of consecutive owned nodes with bounded staged index count
// and bounded UNIQUE neighbour sets; every neighbour reference rewritten as a 16-bit position in its tile's unique
// list; tiles that read no ghost entry listed apart (halo / compute overlap).  Pure host logic, checked by
// verify_tile_plan() in the CPU tests.
struct TileLimits {
  int max_nodes, max_idx, max_uniq, max_puniq;
  int max_vel_blocks = 0;                  // bound on the velocity blocks (node, neighbour node) of one tile; 0 = none
};
struct TilePlan {
  std::vector<int> node_ptr;               // [n_tiles+1]
  std::vector<int> uniq_ptr, uniq_xoff;    // unique neighbour nodes per tile (memory order), x offset of (node, 0)
  std::vector<int> puniq_ptr, puniq_xoff;  // unique neighbour pressure DoFs per tile
  std::vector<unsigned short> nbr_loc;     // parallel to Structure::nbr
  std::vector<unsigned short> pnbr_loc;    // parallel to Structure::pnbr
  std::vector<int> tiles_int, tiles_bnd;   // tiles without / with ghost neighbours
  int n_tiles() const { return (int)node_ptr.size() - 1; }
};
// Returns an empty string on success, else an error message.
std::string build_tile_plan(const Structure& S, const TileLimits& L, TilePlan& P);
// Independent check of every invariant the kernels rely on; returns the number of violations (0 = consistent).
int64_t verify_tile_plan(const Structure& S, const TileLimits& L, const TilePlan& P);

// Streamed velocity operator (velstream.cuh).  The dim x dim blocks (node, neighbour node) of every owned node are grouped
// in QUADS of four consecutive blocks (a node's list is padded to a multiple of four with zero blocks), one lane of a warp
// handles one quad.  Per tile: a 64-byte header; per quad four metadata words (positions of the four neighbours in the
// tile's unique list as 16-bit pairs, index of the quad's node in the tile or 0xffff for padding, first block index within
// the node | real blocks << 16); tiles are padded to NQ = quads rounded up to 32, so every tile's value chunk (dim*dim
// planes of NQ x 4 values) and metadata chunk move with one TMA bulk copy each.  split[] = node-aligned quad offsets that
// divide the tile among the consumer warps.
constexpr int VS_CONSUMERS = 8;                    // consumer warps that share one tile
constexpr int VS_MAX_BLOCKS = 2048;                 // blocks per tile including the padding to quads
constexpr int VS_MAX_QUADS = VS_MAX_BLOCKS / 4;
struct alignas(16) VsTile {
  int n0, nn;                              // first owned node, node count
  int nquad, NQ;                           // quads, quads rounded up to a multiple of 32
  int nq_off;                              // sum of NQ over the preceding tiles (metadata: 4 words per quad; values: dim*dim*4 per quad)
  int u0, nuq;                             // the tile's unique neighbour list in TilePlan::uniq_xoff
  int pad0;
  unsigned short split[VS_CONSUMERS + 1];
  unsigned short pad1[32 - 16 - (VS_CONSUMERS + 1)];
};
static_assert(sizeof(VsTile) == 64, "VsTile is one 64-byte TMA bulk copy");
struct VsPlan {
  std::vector<VsTile> tiles;
  std::vector<uint32_t> meta;              // [4 * sum NQ]
  int64_t total_nq = 0;
};
std::string build_vel_stream(const Structure& S, const TilePlan& P, VsPlan& V);
// independent check (CPU tests): every block of every owned node appears once, in order, with the right neighbour;
// splits are node-aligned and monotone; returns the number of violations
int64_t verify_vel_stream(const Structure& S, const TilePlan& P, const VsPlan& V);

// Coarse level of the two-level velocity preconditioner: the P1 vector space on the mesh vertices (one coarse node per
// pressure DoF, `dim` components each).  `Sc` describes its graph in the same form as the fine level (owned = owned
// vertices, ghosts = ghost vertices in halo order, neighbours = the P1 stencil = pnbr of the vertex node), so the tile /
// stream plans and kernels of the fine level serve it unchanged.  Prolongation P: a vertex node takes its coarse value, a
// line node the mean of its two end vertices; restriction = P^T.
struct CoarseLevel {
  Structure Sc;
  std::vector<int64_t> vedge_ptr;          // [np_own+1]
  std::vector<int> vedge;                  // fine line nodes (local ids, owned or ghost) ending at each owned vertex
  std::vector<int> node_ends;              // [nn_own][2] coarse ids of the end vertices of every owned fine node (equal for vertex nodes)
};
std::string build_coarse(const Structure& S, CoarseLevel& C);

// Expands the local rows to scalar CSR with GLOBAL column indices, rows in local owned
// order (velocity rows then pressure rows); for the bit-exact pattern check.
void export_pattern(const Structure& S, std::vector<int64_t>& rowptr, std::vector<uint32_t>& col);

// Global row index (deal.II numbering) of every local owned row, in local order.
void export_row_gids(const Structure& S, std::vector<int64_t>& gid);

}  // namespace nsb
