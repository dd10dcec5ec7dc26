// TEST INFRASTRUCTURE ONLY (oracle). Never linked or loaded by the product path.
//
// extern "C" shim over the reference's own vendored Clipper 6.4.2, compiled from
// where it lies under /root/reference (see oracle/build_ref.py) into
// oracle/_ref/libclipper_ref.so. It exposes exactly the call sequence that
// DBPostProcessor::UnClip makes
//   (R/pytocr/postprocess/db_postprocess_fast/src/db_postprocess.cpp:41-51):
//   ClipperOffset offset; offset.AddPath(p, jtRound, etClosedPolygon);
//   offset.Execute(soln, distance);
// so the DB oracle (oracle/db_oracle.py) can run the reference's real polygon
// offsetting instead of a restatement of it.
#include "clipper.h"

#include <cstdint>

extern "C" {

// Offsets the closed integer polygon xy[0..2n) by `delta` with round joins.
// Writes the solution paths back to back into out_xy (x,y pairs, int64) and the
// length of each path into out_path_len. Returns the number of paths, or -1 if
// a capacity was exceeded.
int clipper_ref_offset(const int64_t* xy, int n, double delta,
                       int64_t* out_xy, int cap_points,
                       int* out_path_len, int cap_paths) {
  ClipperLib::ClipperOffset offset;
  ClipperLib::Path p;
  for (int i = 0; i < n; ++i)
    p << ClipperLib::IntPoint(xy[2 * i], xy[2 * i + 1]);
  offset.AddPath(p, ClipperLib::jtRound, ClipperLib::etClosedPolygon);
  ClipperLib::Paths soln;
  offset.Execute(soln, delta);
  if ((int)soln.size() > cap_paths) return -1;
  int w = 0;
  for (size_t j = 0; j < soln.size(); ++j) {
    out_path_len[j] = (int)soln[j].size();
    for (size_t i = 0; i < soln[j].size(); ++i) {
      if (w >= cap_points) return -1;
      out_xy[2 * w] = soln[j][i].X;
      out_xy[2 * w + 1] = soln[j][i].Y;
      ++w;
    }
  }
  return (int)soln.size();
}

// The same call sequence for `nquads` quads at once (the timing leg of bench.py: one ctypes call per image
// instead of one per contour). quads: int64 [nquads,4,2]; deltas: double [nquads]. For quad q the points of ALL its
// solution paths are written back to back starting at out_xy[2 * out_start[q]] (out_start has nquads+1 entries) and
// the length of its LAST path goes to last_len[q], the number of its paths to npaths[q] (UnClip's quirky loop,
// db_postprocess.cpp:52-57, needs both). Returns 0, or -1 if cap_points was exceeded.
int clipper_ref_offset_batch(const int64_t* quads, const double* deltas, int nquads,
                             int64_t* out_xy, int cap_points, int* out_start, int* npaths, int* last_len) {
  int w = 0;
  for (int q = 0; q < nquads; ++q) {
    out_start[q] = w;
    ClipperLib::ClipperOffset offset;
    ClipperLib::Path p;
    for (int i = 0; i < 4; ++i) p << ClipperLib::IntPoint(quads[8 * q + 2 * i], quads[8 * q + 2 * i + 1]);
    offset.AddPath(p, ClipperLib::jtRound, ClipperLib::etClosedPolygon);
    ClipperLib::Paths soln;
    offset.Execute(soln, deltas[q]);
    npaths[q] = (int)soln.size();
    last_len[q] = soln.empty() ? 0 : (int)soln.back().size();
    for (size_t j = 0; j < soln.size(); ++j) {
      // the reference copies soln[j][i] for i < soln.back().size() (reads past the end of shorter paths are
      // undefined there; the oracle clamps to the path's own length, as oracle/db_oracle.py does)
      const size_t lim = soln[j].size() < soln.back().size() ? soln[j].size() : soln.back().size();
      for (size_t i = 0; i < lim; ++i) {
        if (w >= cap_points) return -1;
        out_xy[2 * w] = soln[j][i].X;
        out_xy[2 * w + 1] = soln[j][i].Y;
        ++w;
      }
    }
  }
  out_start[nquads] = w;
  return 0;
}

}  // extern "C"
