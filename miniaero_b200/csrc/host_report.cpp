// Mantevo-style YAML run report (host).  The reference carries YAML_Doc / YAML_Element (YAML_Doc.C:27-67,
// YAML_Element.C:97-104) without ever calling them; this emitter keeps their grammar — a two-line header, then
// "key: value" lines indented two spaces per nesting level — and fills it with what the north star asks a run to
// state: cell-updates/s, the HBM-roofline fraction, ranks and blocks, and the setup / run / total times of Main.C.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "host_common.h"
#include "miniaero_b200.h"

namespace {

// one "key: value" node with children (YAML_Element.h:30-84)
class Node {
 public:
  Node(std::string key, std::string value) : key_(std::move(key)), value_(std::move(value)) {}
  Node *add(const std::string &key, const std::string &value = "") {
    kids_.emplace_back(new Node(key, value));
    return kids_.back().get();
  }
  template <class T>
  Node *add_num(const std::string &key, T v) {
    std::ostringstream os;  // default ostream formatting, as YAML_Element::convert_*_to_string
    os << v;
    return add(key, os.str());
  }
  void print(std::string indent, std::string &out) const {  // YAML_Element.C:97-104
    out += indent + key_ + ": " + value_ + "\n";
    indent += "  ";
    for (const auto &k : kids_) k->print(indent, out);
  }
  const std::vector<std::unique_ptr<Node>> &kids() const { return kids_; }

 private:
  std::string key_, value_;
  std::vector<std::unique_ptr<Node>> kids_;
};

// algorithmic bytes per cell-update
double bytes_per_cell_update(const ma_options &o) { return (o.second_order_space || o.viscous) ? 4648.0 : 1928.0; }

}  // namespace

extern "C" int ma_write_yaml_report(const ma_report *r, const char *dir, char *path_out, size_t path_len) {
  if (!r || !r->options || !r->timing) return ma_set_error(MA_ERR_INVALID, "ma_write_yaml_report: null argument");
  const std::string name = r->app_name ? r->app_name : "miniAero-b200";
  const std::string version = r->app_version ? r->app_version : "1.0";
  const ma_options &o = *r->options;
  const ma_timing &t = *r->timing;

  Node root("", "");
  Node *prob = root.add("Problem");
  static const char *kTypes[] = {"3D Sod shock tube", "viscous flat plate", "inviscid ramp"};
  prob->add("type", (o.problem_type >= 0 && o.problem_type <= 2) ? kTypes[o.problem_type] : "unknown");
  prob->add_num("nx", o.nx);
  prob->add_num("ny", o.ny);
  prob->add_num("nz", o.nz);
  prob->add_num("lx", o.lx);
  prob->add_num("ly", o.ly);
  prob->add_num("lz", o.lz);
  prob->add_num("ramp angle", o.angle);
  prob->add_num("time steps", o.ntimesteps);
  prob->add_num("dt", o.dt);
  prob->add("spatial order", o.second_order_space ? "second (Green-Gauss + Venkatakrishnan)" : "first");
  prob->add("viscous", o.viscous ? "yes" : "no");
  prob->add_num("global cells", r->global_cells);

  Node *par = root.add("Parallelism");
  par->add_num("ranks (one per GPU)", r->num_ranks);
  {
    std::ostringstream os;
    os << r->blocks[0] << " x " << r->blocks[1] << " x " << r->blocks[2];
    par->add("blocks", os.str());
  }
  std::string dev = r->device_name ? r->device_name : "";
  if (dev.empty()) {
    int d = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&d) == cudaSuccess && cudaGetDeviceProperties(&p, d) == cudaSuccess) dev = p.name;
  }
  par->add("device", dev.empty() ? "unknown" : dev);
  par->add_num("tiles on this rank", t.num_tiles);
  par->add_num("device bytes on this rank", (long long)t.device_bytes);

  Node *tm = root.add("Timing");
  tm->add_num("setup seconds", r->setup_seconds);
  tm->add_num("run seconds", r->run_seconds);
  tm->add_num("total seconds", r->total_seconds);
  tm->add_num("stepping seconds (device, CUDA events)", t.step_seconds);
  tm->add_num("RK4 steps", t.steps);
  tm->add_num("kernel launches", t.kernel_launches);

  Node *fom = root.add("Figure of merit");
  const double cups = (t.step_seconds > 0 && t.steps > 0) ? (double)r->global_cells * (double)t.steps / t.step_seconds : 0.0;
  fom->add_num("cell-updates per second", cups);
  fom->add_num("algorithmic bytes per cell-update", bytes_per_cell_update(o));
  if (r->hbm_peak_gbs > 0 && r->num_ranks > 0) {
    fom->add_num("HBM peak GB/s (measured, per GPU)", r->hbm_peak_gbs);
    fom->add_num("fraction of HBM roofline", cups / r->num_ranks * bytes_per_cell_update(o) / (r->hbm_peak_gbs * 1e9));
  }

  std::string yaml = "Mini-Application Name: " + name + "\nMini-Application Version: " + version + "\n";  // YAML_Doc.C:29-30
  for (const auto &k : root.kids()) k->print("", yaml);

  time_t raw;
  time(&raw);
  struct tm lt;
  localtime_r(&raw, &lt);
  char stamp[32];
  snprintf(stamp, sizeof(stamp), "%04d:%02d:%02d-%02d:%02d:%02d", lt.tm_year + 1900, lt.tm_mon + 1, lt.tm_mday, lt.tm_hour,
           lt.tm_min, lt.tm_sec);  // YAML_Doc.C:41-42
  const std::string d = (dir && dir[0]) ? dir : ".";
  const std::string path = d + "/" + name + "-" + version + "_" + stamp + ".yaml";  // YAML_Doc.C:45-49
  std::ofstream f(path.c_str());
  if (!f) return ma_set_error(MA_ERR_IO, "cannot open " + path);
  f << yaml;
  f.close();
  if (!f.good()) return ma_set_error(MA_ERR_IO, "write failed: " + path);
  if (path_out && path_len) {
    strncpy(path_out, path.c_str(), path_len - 1);
    path_out[path_len - 1] = '\0';
  }
  return MA_OK;
}
