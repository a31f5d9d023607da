// Stamping a tile pattern into the layout arrays: the O(cells) half of the layout of a structured block, as plain
// functions of (tile, position in the tile's pattern) over a TopoPlan (layout.h).  The same functions run in the CUDA
// kernels of topology_kernels.cu (one CTA per tile) and, on the host, in tools/topology_compare.cpp, which checks them
// against the host builder (layout.cpp) array by array — so the device builder is tested without a GPU as well.
//
// What is reproduced is the loop "tile face lists" of build_layout_impl (layout.cpp) and the renumbering before it;
// reference behaviour behind it: the slot convention of MeshProcessor.C:44,64-118, elem1 = the lower-numbered cell
// (MeshProcessor.C:54-120), boundary sets of Parallel3DMesh.h:382-396.
#pragma once
#include <cstdint>

#include "kernels.h"
#include "mesh_geom.h"

namespace ma {

struct TopoView {
  GridGen g;  // block counts and numbering only (the coordinate tables are not touched)
  int n_owned, n_tiles, slot_stride, halo_stride, import_capacity;
  int bc_of_face[6];
  const TileInfoDev *tiles;
  const int *tile_pattern;          // [n_tiles]
  const int *tile_origin;           // [n_tiles][3]
  const int *tile_nb;               // [n_tiles][6]
  const unsigned char *tile_launch; // [n_tiles]
  // patterns, flattened
  const int *pat_ext;       // [npat][3]
  const int *pat_dummy;     // [npat] face index of the padding duplicate, -1
  const int *pat_cell_off;  // [npat] first entry in cell_abc
  const int *pat_rank_off;  // [npat] first entry in rank_of
  const int *pat_face_off;  // [npat] first entry in face_lc / face_slot
  const uint32_t *cell_abc;
  const uint16_t *rank_of;
  const uint16_t *face_lc;
  const uint8_t *face_slot;
  // outputs (device layout arrays, see layout.h / kernels.h)
  int *new2old, *old2new;
  uint16_t *slot_face, *slot_nbr;
  uint32_t *face_lr, *face_code;
  int *tile_halo, *tile_pub;
};

// renumbered id of the owned block cell (i, j, k)
MA_HD inline int topo_new_id(const TopoView &t, int tile, int i, int j, int k) {
  const int p = t.tile_pattern[tile];
  const int *o = t.tile_origin + 3 * tile, *e = t.pat_ext + 3 * p;
  const int a = i - o[0], b = j - o[1], c = k - o[2];
  return t.tiles[tile].cell_start + t.rank_of[t.pat_rank_off[p] + (a * e[1] + b) * e[2] + c];
}

// cell lc of tile k: the renumbering maps
MA_HD inline void topo_stamp_cell(const TopoView &t, int k, int lc) {
  const int p = t.tile_pattern[k];
  const uint32_t abc = t.cell_abc[t.pat_cell_off[p] + lc];
  const int *o = t.tile_origin + 3 * k;
  const int old = (int)t.g.cell_id(o[0] + (int)(abc & 255u), o[1] + (int)((abc >> 8) & 255u), o[2] + (int)(abc >> 16));
  const int nw = t.tiles[k].cell_start + lc;
  t.new2old[nw] = old;
  t.old2new[old] = nw;
}

// face e of tile k: connectivity, slot maps, face code, outside-cell list
MA_HD inline void topo_stamp_face(const TopoView &t, int k, int e) {
  const TileInfoDev T = t.tiles[k];
  const int p = t.tile_pattern[k];
  const int lc = t.face_lc[t.pat_face_off[p] + e], s = t.face_slot[t.pat_face_off[p] + e];
  const bool dummy = e == t.pat_dummy[p];
  const bool cut = e >= T.cut_start;
  const uint32_t abc = t.cell_abc[t.pat_cell_off[p] + lc];
  const int *o = t.tile_origin + 3 * k, *ext = t.pat_ext + 3 * p;
  const int a = (int)(abc & 255u), b = (int)((abc >> 8) & 255u), cc = (int)(abc >> 16);
  const int i = o[0] + a, j = o[1] + b, kk = o[2] + cc;
  int di, dj, dk;
  face_dir(s, di, dj, dk);
  const long c_old = t.g.cell_id(i, j, kk);
  const long nb_old = t.g.cell_id(i + di, j + dj, kk + dk);
  const int c = T.cell_start + lc;
  const int shift = T.cell_start & 1, halo_base = (shift + T.cell_count + 1) & ~1;
  const size_t jf = (size_t)T.face_start + e;
  // elem1 = the lower-numbered of the two cells (the only cell of a boundary face); the face code names it
  {
    int ci = i, cj = j, ck = kk, f = s;
    if (nb_old >= 0 && nb_old < c_old) ci += di, cj += dj, ck += dk, f = opposite_face(s);
    const long lat = ((long)(ci + 1) * (t.g.b.n[1] + 2) + (cj + 1)) * (t.g.b.n[2] + 2) + (ck + 1);
    t.face_code[jf] = (uint32_t)(lat * 8 + f);
  }
  if (nb_old < 0) {  // boundary face
    const int bc = t.bc_of_face[s];
    if (!dummy) t.slot_face[(size_t)s * t.slot_stride + c] = (uint16_t)(e | (1 << 14));
    t.face_lr[jf] = (uint32_t)(shift + lc) | ((uint32_t)(0xFFFF - bc) << 16);
    return;
  }
  const int side = nb_old > c_old ? 0 : 1;
  if (!dummy) t.slot_face[(size_t)s * t.slot_stride + c] = (uint16_t)(e | (side << 15));
  int oth_local;
  if (cut) {
    int oth_new;
    if (nb_old >= t.n_owned) {
      oth_new = (int)nb_old;  // ghosts keep their ids
    } else {
      oth_new = topo_new_id(t, t.tile_nb[6 * k + s], i + di, j + dj, kk + dk);
    }
    oth_local = halo_base + (e - T.cut_start);
    t.tile_halo[(size_t)k * t.halo_stride + (e - T.cut_start)] = oth_new;
  } else {
    const int olc = t.rank_of[t.pat_rank_off[p] + ((a + di) * ext[1] + (b + dj)) * ext[2] + (cc + dk)];
    const int oth_new = T.cell_start + olc;
    oth_local = shift + olc;
    if (!dummy) {
      const int os = opposite_face(s);
      t.slot_face[(size_t)os * t.slot_stride + oth_new] = (uint16_t)(e | ((1 - side) << 15));
      t.slot_nbr[(size_t)os * t.slot_stride + oth_new] = (uint16_t)(shift + lc);
    }
  }
  if (!dummy) t.slot_nbr[(size_t)s * t.slot_stride + c] = (uint16_t)oth_local;
  t.face_lr[jf] = side == 0 ? ((uint32_t)(shift + lc) | ((uint32_t)oth_local << 16))
                            : ((uint32_t)oth_local | ((uint32_t)(shift + lc) << 16));
}

// evaluated cut face cut_start + q of tile k: where it publishes its flux (after every tile's faces are stamped)
MA_HD inline void topo_stamp_pub(const TopoView &t, int k, int q) {
  const TileInfoDev T = t.tiles[k];
  const int e = T.cut_start + q;
  const int p = t.tile_pattern[k];
  if (e >= T.n_eval || e == t.pat_dummy[p]) return;
  const int s = t.face_slot[t.pat_face_off[p] + e];
  const int k2 = t.tile_nb[6 * k + s];
  if (k2 < 0 || !(t.tile_launch[k] < t.tile_launch[k2])) return;  // a ghost, or the other tile evaluates the face itself
  const int on = t.tile_halo[(size_t)k * t.halo_stride + q];
  const TileInfoDev T2 = t.tiles[k2];
  const int e2 = t.slot_face[(size_t)opposite_face(s) * t.slot_stride + on] & 0x3fff;
  t.tile_pub[(size_t)k * t.halo_stride + q] = T2.imp_area * 5 * t.import_capacity + (e2 - T2.n_eval);
}

}  // namespace ma
