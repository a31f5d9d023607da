// Host-side construction of the device data layout: cell renumbering into spatial tiles,
// tile-packed structure-of-arrays face lists, cell -> tile-face slot maps.
//
// This replaces the reference's Faces<Device>/Cells<Device> containers (Faces.h:42-71, Cells.h:41-76),
// its BinSort face permutation (Faces.h:128-144) and its 6-slot scratch arrays cell_flux_ /
// cell_gradient_ / stored_{min,max,limiter}_ (Cells.h:70-71, StencilLimiter.h:525-527).
#pragma once
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

#include "host_common.h"
#include "mesh_geom.h"
#include "miniaero_b200.h"

namespace ma {

// Boundary faces are encoded in the tile-face "right cell" entry as -1 - ma_bc_type.
inline int bc_code(int type) { return -1 - type; }

struct TileInfo {
  int cell_start;  // first (renumbered) cell of the tile
  int cell_count;
  int face_start;  // first entry of the tile in the tile-packed face arrays
  int face_count;
  int cut_start;   // tile-local index of the first "cut" face (one cell in the tile, the other — owned by a
                   // neighbouring tile or a ghost — outside); faces [0, cut_start) are closed or boundary faces
  int halo_start;  // first entry of the tile in tile_halo; cut face e's outside cell is tile_halo[halo_start + e - cut_start]
  // Shared cut faces (HostLayout::share_cut_faces): a face between two tiles is evaluated by ONE of them — the one whose
  // flux launch comes first — which also publishes the flux to the other tile's import area; without sharing every
  // cut face is evaluated by both tiles.  Faces [0, n_eval) are evaluated by this tile's CTA (closed / boundary faces
  // [0, cut_start), then the cut faces it evaluates); faces [n_eval, face_count) are cut faces whose flux is imported.
  // n_eval is even when the tile imports anything (a duplicate of an evaluated face pads an odd count), so that the
  // imported columns start 16-byte aligned in the staged flux array.
  int n_eval;
  int imp_area;    // index of the tile's import area in the cut-flux exchange buffer, -1: none
};

struct HostLayout {
  int n_owned = 0, n_ghost = 0;
  int stride = 0;  // SoA component stride of cell arrays (>= n_owned + n_ghost, multiple of 32)
  int tile_dims[3] = {0, 0, 0};
  int max_tile_cells = 0, max_tile_faces = 0;  // max_tile_cells: the requested brick size (upper bound)
  int max_tile_cells_real = 0;                 // largest tile actually built
  int n_tiles = 0;
  int n_interior_tiles = 0;  // tiles [0, n_interior_tiles) touch no ghost cell
  long n_tile_faces = 0;     // length of the tile-packed face arrays (with per-tile padding)
  long n_tile_faces_real = 0;

  BigVec<int> new2old, old2new;  // owned + ghost cells
  std::vector<TileInfo> tiles;

  // cell SoA (renumbered): xyz[3][stride], vol[stride]
  BigVec<double> cell_xyz, cell_vol;
  // slot map: slot_face[s][cell] (owned cells only, stride = n_owned rounded up to 32):
  // bits 0..13 tile-local face index, bit 14 = 1 for a boundary face, bit 15 = 1 when the cell is elem2 (right)
  BigVec<uint16_t> slot_face;
  int slot_stride = 0;

  // neighbour map (FAST kernels): slot_nbr[s][cell] = position of the cell across slot s in the tile's staged cell
  // list (see face_lr), 0xFFFF for a boundary face
  BigVec<uint16_t> slot_nbr;

  // tile-packed face geometry.
  // with_tangents (STRICT arithmetic): global SoA, component g of tile face j at geom[g*n_tile_faces + j],
  //                g = 0-2 normal, 3-5 tangent, 6-8 binormal, 9-11 centroid;
  // otherwise (FAST arithmetic, never reads tangent / binormal): tile-blocked SoA, so that one contiguous run holds
  //                everything a tile needs: component g of tile T's face e at geom[6*T.face_start + g*fcp + e] with
  //                fcp = T.face_count rounded up to 16, g = 0-2 normal, 3-5 centroid
  int geom_components = 12;
  double max_frame_error = 0.0;  // worst deviation of (n^, t, b/|a|) from an orthonormal frame over all faces
  BigVec<double> face_geom;
  // structured path with device-side geometry (build_layout_structured, defer_geometry): face_geom, cell_xyz and
  // cell_vol stay empty; face_code[j] = (elem1 cell in the block's (n+2)^3 lattice) * 8 + elem1 local face says which
  // face tile face j is, and new2old which cell, for geom_kernels.cu to evaluate mesh_geom.h on the device
  bool geometry_deferred = false;
  BigVec<uint32_t> face_code;
  BigVec<int> face_left, face_right;  // renumbered cell ids; right < 0 -> boundary code
  // tile-local connectivity: low 16 bits = left cell, high 16 bits = right cell, as POSITIONS in the tile's staged
  // cell list: own cell lc sits at (cell_start & 1) + lc (the staging copy starts at the even cell below
  // cell_start: 16-byte alignment of the bulk copies), the outside cell of cut face e at
  // halo_base + (e - cut_start) with halo_base = ((cell_start & 1) + cell_count) rounded up to even;
  // a boundary face has right = 0xFFFF - ma_bc_type
  BigVec<uint32_t> face_lr;
  BigVec<int> tile_halo;  // renumbered id of the outside cell of every cut face, halo_stride entries per tile
  int max_tile_local = 0;      // largest staged cell list of a tile (halo_base + cut faces)
  int max_tile_halo = 0;       // largest number of cut faces of a tile
  int halo_stride = 0;         // tile k's outside cells are tile_halo[k * halo_stride ...], padded with -1

  // ---- shared cut faces (FAST staged flux kernel; see TileInfo)
  bool share_cut_faces = false;
  // tiles per flux launch, in tile order: [interior first pass, interior second pass, boundary first pass, boundary
  // second pass] (boundary = touches a ghost cell).  The passes are the two colours of the tile lattice's checkerboard:
  // face-adjacent bricks never share a launch, and the owner of a cut face is the tile of the earlier launch (tiles of
  // one launch that do touch — irregular meshes — both evaluate the face, as without sharing).
  int launch_count[4] = {0, 0, 0, 0};
  int import_capacity = 0;  // entries per component of an import area (even); component k of import j at 5*cap*area + k*cap + j
  int n_import_areas = 0;
  // [n_tiles][halo_stride], parallel to tile_halo: where evaluated cut face cut_start + q publishes its flux (index of
  // component 0 in the exchange buffer), -1: nowhere (the other side is a ghost, or evaluates the face itself)
  BigVec<int> tile_pub;

  // true: only the O(tiles) members and the exchange lists below are filled; the O(cells) arrays are built on the
  // device from a TopoPlan (build_topology_plan)
  bool topology_on_device = false;

  // halo lists in renumbered ids, grouped by neighbour rank ascending
  std::vector<int> send_ids, recv_ids;
  std::vector<int> peer_rank, peer_send_count, peer_recv_count;
};

// Returns MA_OK or sets the error text.  tile_dims: requested cells per tile per direction.
int build_layout(const ma_mesh &mesh, const int tile_dims[3], bool with_tangents, HostLayout &L,
                 bool share_cut_faces = false);

// The same layout for one block of the in-code structured mesh (the reference's Parallel3DMesh / MeshProcessor,
// Parallel3DMesh.h:173-449) WITHOUT materialising the reference-format arrays: connectivity, numbering and exchange
// lists follow from (i, j, k) (mesh_geom.h).  Bit for bit the layout build_layout() makes of ma_mesh_generate()'s
// mesh (tests/test_layout.py).  defer_geometry: leave face / cell geometry to the device (see HostLayout).
struct StructuredGrid {
  GridGen gen;
  GridTables tables;
};
int build_layout_structured(const ma_options &opt, int rank, int num_ranks, const int tile_dims[3], bool with_tangents,
                            bool defer_geometry, HostLayout &L, StructuredGrid *grid, bool share_cut_faces = false);

// ---- topology on the device (structured blocks, FAST staged kernels) ------------------------------------------------
// The O(cells) part of the layout — cell renumbering, slot maps, tile-local connectivity, face codes, outside-cell
// lists, publish slots — is a pure function of (tile, position in the tile's pattern): tiles of a structured block
// fall into a few dozen PATTERNS (brick extents x what lies behind each of the six sides x parity of the first cell),
// and a pattern fixes the tile-local order of cells and faces.  build_topology_plan() does the O(tiles) and
// O(patterns) work on the host — tile order, descriptors, one TileOrder per pattern (the same compute_tile_order the
// host builder uses, so the two builders cannot drift apart) — and topology_kernels.cu stamps the patterns on the
// device.  HostLayout then carries only its O(tiles) members and the (renumbered) exchange lists; the O(cells)
// vectors stay empty.
struct TopoPattern {
  int ext[3];                       // brick extents (cells)
  int cell_count, face_count, cut_start, n_eval;
  int dummy_face;                   // face index of the padding duplicate, -1: none
  std::vector<uint32_t> cell_abc;   // [cell_count] brick-local (a, b, c) of the lc-th cell: a | b << 8 | c << 16
  std::vector<uint16_t> rank_of;    // [ext0 * ext1 * ext2] (a * ext1 + b) * ext2 + c -> lc
  std::vector<uint16_t> face_lc;    // [face_count] emitting cell of face e
  std::vector<uint8_t> face_slot;   // [face_count] its slot
};
struct TopoPlan {
  std::vector<TopoPattern> patterns;
  std::vector<int> tile_pattern;    // [n_tiles]
  std::vector<int> tile_origin;     // [n_tiles][3] block-local (i, j, k) of the brick's first cell
  std::vector<int> tile_nb;         // [n_tiles][6] tile behind slot direction s; -1 domain boundary, -2 ghost layer
  std::vector<uint8_t> tile_launch; // [n_tiles] flux launch class (layout.h: HostLayout::launch_count)
  int bc_of_face[6];                // ma_bc_type of the domain side behind local face f
};
int build_topology_plan(const ma_options &opt, int rank, int num_ranks, const int tile_dims[3], bool share_cut_faces,
                        HostLayout &L, StructuredGrid *grid, TopoPlan &plan);

}  // namespace ma
