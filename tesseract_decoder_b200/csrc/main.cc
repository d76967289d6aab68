// `tesseract` CLI over the C ABI: the decode-path flags of the reference CLI (tesseract_main.cc:350-538) with the
// same names and defaults, the same accounting (tesseract_main.cc:592-616: a low-confidence shot is not a logical
// error) and the same final line / --stats-out keys (tesseract_main.cc:657-699).  Shots are decoded in batches on
// the GPU(s) instead of per-thread on the CPU; --threads is accepted and ignored, --gpus N shards batches over N
// devices (one handle each).  stim is not available here, so --circuit goes through the built-in circuit->DEM
// analyzer and --sample-num-shots through the i.i.d. DEM sampler; shot files are b8 or 01.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <sstream>
#include <thread>

#include "decoder.h"

using namespace tesseract_b200;

namespace {

struct Args {
  std::string circuit_path, dem_path, in_fname, in_format = "b8", obs_in_fname, obs_in_format = "b8", out_fname,
      out_format = "b8", stats_out_fname, dem_out_fname;
  bool in_includes_appended_observables = false;
  size_t num_det_orders = 1;              // tesseract_main.cc:358
  int det_order_method = 1;               // DetIndex (tesseract_main.cc:230-237)
  uint64_t det_order_seed = 518278944;    // tesseract_main.cc:379
  uint64_t sample_num_shots = 0, sample_seed = std::random_device{}();
  uint64_t max_errors = UINT64_MAX, shot_range_begin = 0, shot_range_end = 0;
  int det_beam = INF_DET_BEAM;            // tesseract_main.cc:482
  double det_penalty = 0;
  bool beam_climbing = false, no_revisit_dets = false, at_most_two = false, no_merge_errors = false, verbose = false,
       print_stats = false, convert_only = false;
  size_t pqlimit = std::numeric_limits<size_t>::max();  // tesseract_main.cc:503
  int num_threads = (int)std::thread::hardware_concurrency();
  int gpus = 1;
  uint64_t batch = 1 << 18;
};

[[noreturn]] void usage(const std::string& why) {
  std::cerr << why << "\n"
            << "usage: tesseract (--dem FILE | --circuit FILE) (--sample-num-shots N [--sample-seed S] | --in FILE [--in-format b8|01]\n"
               "       [--in-includes-appended-observables] [--obs_in FILE [--obs-in-format b8|01]]) [--out FILE [--out-format b8|01]]\n"
               "       [--beam N] [--det-penalty X] [--beam-climbing] [--no-revisit-dets] [--at-most-two-errors-per-detector]\n"
               "       [--pqlimit N] [--num-det-orders N] [--det-order-seed S] [--det-order-bfs|--det-order-index|--det-order-coordinate]\n"
               "       [--no-merge-errors] [--max-errors N] [--shot-range-begin A --shot-range-end B] [--stats-out FILE] [--dem-out FILE]\n"
               "       [--print-stats]\n"
               "       [--gpus N] [--threads N (ignored)] [--verbose]\n";
  std::exit(EXIT_FAILURE);
}

std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::invalid_argument("Failed to open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

// A bit index in a text shot file: decimal digits only (std::stoull would take "-3" or "12abc").
uint64_t parse_index(std::string tok, const std::string& path) {
  while (!tok.empty() && (tok.back() == ' ' || tok.back() == '\t')) tok.pop_back();
  tok.erase(0, tok.find_first_not_of(" \t") == std::string::npos ? tok.size() : tok.find_first_not_of(" \t"));
  if (tok.empty() || tok.size() > 19 || tok.find_first_not_of("0123456789") != std::string::npos)
    throw std::invalid_argument(path + ": '" + tok + "' is not a bit index");
  return std::stoull(tok);
}

// Reads shots as rows of `bits` bits into a [shots][ceil(bits/8)] b8 matrix.
// (`dets_prefix`: in the dets format, L# tokens index bits after the first dets_prefix bits.)
std::vector<uint8_t> read_shots(const std::string& path, const std::string& format, uint64_t bits, uint64_t& shots_out,
                                uint64_t dets_prefix = 0) {
  const std::string raw = read_file(path);
  const uint64_t nb = (bits + 7) / 8;
  std::vector<uint8_t> out;
  if (format == "b8") {
    if (nb == 0 || raw.size() % nb) throw std::invalid_argument(path + ": size is not a multiple of the b8 shot size");
    out.assign(raw.begin(), raw.end());
    shots_out = raw.size() / nb;
  } else if (format == "01") {
    shots_out = 0;
    std::istringstream ss(raw);
    std::string line;
    while (std::getline(ss, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.empty()) continue;
      if (line.size() != bits) throw std::invalid_argument(path + ": a 01 line does not have " + std::to_string(bits) + " bits");
      out.resize(out.size() + nb, 0);
      uint8_t* row = out.data() + out.size() - nb;
      for (uint64_t i = 0; i < bits; ++i) {
        if (line[i] == '1') row[i / 8] |= (uint8_t)(1u << (i % 8));
        else if (line[i] != '0') throw std::invalid_argument(path + ": unexpected character in 01 data");
      }
      ++shots_out;
    }
  } else if (format == "hits") {  // one line per shot: comma-separated indices of the set bits
    shots_out = 0;
    std::istringstream ss(raw);
    std::string line;
    while (std::getline(ss, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      out.resize(out.size() + nb, 0);
      uint8_t* row = out.data() + out.size() - nb;
      std::istringstream ls(line);
      std::string tok;
      while (std::getline(ls, tok, ',')) {
        if (tok.empty()) continue;
        const uint64_t i = parse_index(tok, path);
        if (i >= bits) throw std::invalid_argument(path + ": hit index " + tok + " is out of range");
        row[i / 8] ^= (uint8_t)(1u << (i % 8));
      }
      ++shots_out;
    }
  } else if (format == "dets") {  // "shot D3 D7 L0": D# are the first num_detectors bits, L# follow them
    shots_out = 0;
    std::istringstream ss(raw);
    std::string line;
    while (std::getline(ss, line)) {
      std::istringstream ls(line);
      std::string tok;
      if (!(ls >> tok)) continue;
      if (tok != "shot") throw std::invalid_argument(path + ": a dets line must start with 'shot'");
      out.resize(out.size() + nb, 0);
      uint8_t* row = out.data() + out.size() - nb;
      while (ls >> tok) {
        if (tok.size() < 2 || (tok[0] != 'D' && tok[0] != 'L' && tok[0] != 'M')) throw std::invalid_argument(path + ": bad dets token " + tok);
        uint64_t i = parse_index(tok.substr(1), path);
        if (tok[0] == 'L') i += dets_prefix;
        if (i >= bits) throw std::invalid_argument(path + ": dets token " + tok + " is out of range");
        row[i / 8] ^= (uint8_t)(1u << (i % 8));
      }
      ++shots_out;
    }
  } else if (format == "r8") {  // run lengths of zeros before each one; 0xFF = 255 zeros, no one; a one is appended per shot
    shots_out = 0;
    size_t at = 0;
    while (at < raw.size()) {
      out.resize(out.size() + nb, 0);
      uint8_t* row = out.data() + out.size() - nb;
      uint64_t pos = 0;
      while (true) {
        if (at >= raw.size()) throw std::invalid_argument(path + ": r8 data ends inside a shot");
        const uint8_t run = (uint8_t)raw[at++];
        pos += run;
        if (run == 0xFF) continue;
        if (pos == bits) break;  // the terminating one
        if (pos > bits) throw std::invalid_argument(path + ": r8 run passes the end of a shot");
        row[pos / 8] |= (uint8_t)(1u << (pos % 8));
        ++pos;
      }
      ++shots_out;
    }
  } else if (format == "ptb64") {
    // Partially transposed, bit-packed in groups of 64 shots: per group, for each bit index k, one little-endian
    // 64-bit word whose bit s is bit k of the group's shot s.  The shot count is a multiple of 64.
    if (bits == 0 || raw.size() % (bits * 8)) throw std::invalid_argument(path + ": size is not a multiple of the ptb64 group size (64 shots)");
    const uint64_t groups = raw.size() / (bits * 8);
    shots_out = groups * 64;
    out.assign((size_t)(shots_out * nb), 0);
    for (uint64_t g = 0; g < groups; ++g)
      for (uint64_t k = 0; k < bits; ++k) {
        uint64_t word = 0;
        for (int b = 0; b < 8; ++b) word |= (uint64_t)(uint8_t)raw[(size_t)((g * bits + k) * 8 + b)] << (8 * b);
        for (; word; word &= word - 1) {
          const uint64_t sh = g * 64 + (uint64_t)__builtin_ctzll(word);
          out[(size_t)(sh * nb + k / 8)] |= (uint8_t)(1u << (k % 8));
        }
      }
  } else {
    throw std::invalid_argument("unsupported shot format '" + format + "' (b8, 01, hits, dets, r8 and ptb64 are supported)");
  }
  return out;
}

// Writes one shot of `bits` bits (stim result formats b8, 01, hits, dets with the given token letter, r8).
void write_shot(std::ostream& out, const std::string& format, const uint8_t* row, uint64_t bits, char letter) {
  auto bit = [&](uint64_t i) { return (row[i / 8] >> (i % 8)) & 1; };
  if (format == "b8") {
    out.write(reinterpret_cast<const char*>(row), (std::streamsize)((bits + 7) / 8));
  } else if (format == "01") {
    for (uint64_t i = 0; i < bits; ++i) out.put(bit(i) ? '1' : '0');
    out.put('\n');
  } else if (format == "hits") {
    bool first = true;
    for (uint64_t i = 0; i < bits; ++i)
      if (bit(i)) {
        if (!first) out.put(',');
        out << i;
        first = false;
      }
    out.put('\n');
  } else if (format == "dets") {
    out << "shot";
    for (uint64_t i = 0; i < bits; ++i)
      if (bit(i)) out << ' ' << letter << i;
    out.put('\n');
  } else if (format == "r8") {
    uint64_t run = 0;
    for (uint64_t i = 0; i <= bits; ++i) {
      if (i == bits || bit(i)) {
        while (run >= 255) {
          out.put((char)0xFF);
          run -= 255;
        }
        out.put((char)run);
        run = 0;
      } else {
        ++run;
      }
    }
  } else {
    throw std::invalid_argument("unsupported --out-format '" + format + "' (b8, 01, hits, dets, r8 and ptb64 are supported)");
  }
}

// Shot-by-shot writer.  ptb64 transposes groups of 64 shots, so its rows are held back until a group is complete;
// finish() refuses a trailing partial group (the format cannot represent it).
struct ShotWriter {
  std::ostream& out;
  std::string format;
  uint64_t bits;
  char letter;
  std::vector<uint8_t> group;  // ptb64: up to 64 pending rows
  uint64_t pending = 0;
  ShotWriter(std::ostream& o, std::string f, uint64_t b, char l) : out(o), format(std::move(f)), bits(b), letter(l) {}
  void write(const uint8_t* row) {
    if (format != "ptb64") return write_shot(out, format, row, bits, letter);
    const uint64_t nb = (bits + 7) / 8;
    group.insert(group.end(), row, row + nb);
    if (++pending < 64) return;
    for (uint64_t k = 0; k < bits; ++k) {
      uint64_t word = 0;
      for (uint64_t sh = 0; sh < 64; ++sh) word |= (uint64_t)((group[(size_t)(sh * nb + k / 8)] >> (k % 8)) & 1) << sh;
      for (int b = 0; b < 8; ++b) out.put((char)((word >> (8 * b)) & 0xFF));
    }
    group.clear();
    pending = 0;
  }
  void finish() {
    if (format == "ptb64" && pending) throw std::invalid_argument("ptb64 output needs a multiple of 64 shots");
  }
};

}  // namespace

int main(int argc, char** argv) {
  Args a;
  std::map<std::string, std::string*> str_opts = {
      {"--circuit", &a.circuit_path}, {"--dem", &a.dem_path}, {"--in", &a.in_fname}, {"--in-format", &a.in_format},
      {"--in_format", &a.in_format}, {"--obs_in", &a.obs_in_fname}, {"--obs-in", &a.obs_in_fname},
      {"--obs-in-format", &a.obs_in_format}, {"--obs_in_format", &a.obs_in_format}, {"--out", &a.out_fname},
      {"--out-format", &a.out_format}, {"--stats-out", &a.stats_out_fname}, {"--dem-out", &a.dem_out_fname}};
  std::map<std::string, bool*> flag_opts = {
      {"--beam-climbing", &a.beam_climbing}, {"--no-revisit-dets", &a.no_revisit_dets},
      {"--at-most-two-errors-per-detector", &a.at_most_two}, {"--no-merge-errors", &a.no_merge_errors},
      {"--verbose", &a.verbose}, {"--print-stats", &a.print_stats}, {"--convert-only", &a.convert_only},
      {"--in-includes-appended-observables", &a.in_includes_appended_observables},
      {"--in_includes_appended_observables", &a.in_includes_appended_observables}};
  TesseractConfig config_sparsify;  // carries the --sparsify-* flags
  bool has_base = false, has_max = false, has_limit = false;
  int det_order_flags = 0;
  bool threads_given = false;
  try {
    for (int i = 1; i < argc; ++i) {
      const std::string k = argv[i];
      auto need = [&]() -> std::string {
        if (i + 1 >= argc) usage(k + " needs a value");
        return argv[++i];
      };
      // numbers: the whole token must parse, and the message names the flag (std::stoull alone says "stoull")
      auto u64 = [&]() -> uint64_t {
        const std::string v = need();
        if (v.empty() || v.size() > 20 || v.find_first_not_of("0123456789") != std::string::npos)
          throw std::invalid_argument(k + ": '" + v + "' is not an unsigned integer");
        try {
          return std::stoull(v);
        } catch (const std::exception&) {
          throw std::invalid_argument(k + ": '" + v + "' is out of range");
        }
      };
      auto i32 = [&]() -> int {
        const std::string v = need();
        size_t used = 0;
        try {
          const int r = std::stoi(v, &used);
          if (used == v.size()) return r;
        } catch (const std::exception&) {
        }
        throw std::invalid_argument(k + ": '" + v + "' is not an integer");
      };
      auto f64 = [&]() -> double {
        const std::string v = need();
        size_t used = 0;
        try {
          const double r = std::stod(v, &used);
          if (used == v.size()) return r;
        } catch (const std::exception&) {
        }
        throw std::invalid_argument(k + ": '" + v + "' is not a number");
      };
      if (str_opts.count(k)) *str_opts[k] = need();
      else if (flag_opts.count(k)) *flag_opts[k] = true;
      else if (k == "--num-det-orders") a.num_det_orders = u64();
      else if (k == "--det-order-bfs") a.det_order_method = 0, ++det_order_flags;
      else if (k == "--det-order-index") a.det_order_method = 1, ++det_order_flags;
      else if (k == "--det-order-coordinate") a.det_order_method = 2, ++det_order_flags;
      else if (k == "--det-order-seed") a.det_order_seed = u64();
      else if (k == "--sample-num-shots") a.sample_num_shots = u64();
      else if (k == "--sample-seed") a.sample_seed = u64();
      else if (k == "--max-errors") a.max_errors = u64();
      else if (k == "--shot-range-begin") a.shot_range_begin = u64();
      else if (k == "--shot-range-end") a.shot_range_end = u64();
      else if (k == "--beam") a.det_beam = i32();
      else if (k == "--det-penalty") a.det_penalty = f64();
      else if (k == "--pqlimit") a.pqlimit = u64();
      else if (k == "--threads") a.num_threads = i32(), threads_given = true;
      else if (k == "--gpus") a.gpus = i32();
      else if (k == "--batch") a.batch = u64();
      else if (k == "--sparsify-errors") config_sparsify.sparsify_errors = true;
      else if (k == "--sparsify-base-degree") config_sparsify.sparsify_base_degree = i32(), has_base = true;
      else if (k == "--sparsify-max-degree") config_sparsify.sparsify_max_degree = i32(), has_max = true;
      else if (k == "--sparsify-reactivate-limit") config_sparsify.sparsify_reactivate_limit = i32(), has_limit = true;
      else if (k == "-h" || k == "--help") usage("tesseract (B200): A* most-likely-error decoder");
      else usage("unknown argument " + k);
    }
    // Args::validate (tesseract_main.cc:98-180)
    if (a.circuit_path.empty() == a.dem_path.empty()) usage("Must provide exactly one of --circuit or --dem");
    if ((a.sample_num_shots != 0) == !a.in_fname.empty()) usage("Must provide exactly one of --sample-num-shots or --in");
    if (det_order_flags > 1)
      throw std::invalid_argument("Only one of --det-order-bfs, --det-order-index, or --det-order-coordinate may be set.");
    for (const std::string* f : {&a.in_format, &a.obs_in_format, &a.out_format})
      if (*f != "01" && *f != "b8" && *f != "r8" && *f != "hits" && *f != "dets" && *f != "ptb64")
        throw std::invalid_argument("Invalid format: " + *f);
    if (!a.obs_in_fname.empty() && a.in_fname.empty())
      throw std::invalid_argument("Cannot load observable flips without a corresponding detection event data file.");
    if (threads_given && a.num_threads < 1) throw std::invalid_argument("--threads must be at least 1.");  // accepted and ignored otherwise
    if ((a.shot_range_begin || a.shot_range_end) && a.shot_range_end < a.shot_range_begin)
      throw std::invalid_argument("Provided shot range must have end >= begin.");
    if (a.beam_climbing && a.det_beam == INF_DET_BEAM) usage("Beam climbing requires a finite beam");
    if (a.gpus < 1) usage("--gpus must be >= 1");
    if (!config_sparsify.sparsify_errors) {  // tesseract_main.cc:155-179
      if (has_base || has_max || has_limit)
        throw std::invalid_argument(
            "Cannot use --sparsify-base-degree, --sparsify-max-degree, or --sparsify-reactivate-limit without --sparsify-errors");
    } else {
      if (!has_base) throw std::invalid_argument("Must specify --sparsify-base-degree when --sparsify-errors is enabled.");
      if (config_sparsify.sparsify_base_degree <= 0) throw std::invalid_argument("--sparsify-base-degree must be > 0.");
      if (has_limit && config_sparsify.sparsify_reactivate_limit < -1)
        throw std::invalid_argument("--sparsify-reactivate-limit must be >= -1.");
      if (has_max && config_sparsify.sparsify_max_degree < config_sparsify.sparsify_base_degree)
        throw std::invalid_argument("--sparsify-max-degree must be >= --sparsify-base-degree.");
    }

    if (a.convert_only) {
      // Host-only helper: re-encode the detection-event file given by --in into --out / --out-format (no decoding).
      if (a.in_fname.empty() || a.out_fname.empty() || a.dem_path.empty()) usage("--convert-only needs --dem, --in and --out");
      const uint64_t bits = tsr_dem_count_detectors(read_file(a.dem_path).c_str());
      if (bits == UINT64_MAX) throw std::invalid_argument(tsr_last_error());
      uint64_t n = 0;
      const std::vector<uint8_t> rows = read_shots(a.in_fname, a.in_format, bits, n);
      std::ofstream f(a.out_fname, std::ios::binary);
      if (!f) throw std::invalid_argument("Failed to open " + a.out_fname);
      ShotWriter w(f, a.out_format, bits, 'D');
      for (uint64_t k = 0; k < n; ++k) w.write(rows.data() + k * ((bits + 7) / 8));
      w.finish();
      return EXIT_SUCCESS;
    }
    std::string dem_text;
    if (!a.circuit_path.empty()) {
      char* out = nullptr;
      check(tsr_circuit_to_dem(read_file(a.circuit_path).c_str(), &out));
      dem_text = out;
      tsr_free(out);
    } else {
      dem_text = read_file(a.dem_path);
    }

    TesseractConfig config;
    config.dem = dem_text;
    config.det_beam = a.det_beam;
    config.det_penalty = a.det_penalty;
    config.beam_climbing = a.beam_climbing;
    config.no_revisit_dets = a.no_revisit_dets;
    config.at_most_two_errors_per_detector = a.at_most_two;
    config.merge_errors = !a.no_merge_errors;
    config.pqlimit = a.pqlimit;
    config.verbose = a.verbose;
    config.det_orders = build_det_orders(dem_text, a.num_det_orders, (DetOrder)a.det_order_method, a.det_order_seed);
    config.sparsify_errors = config_sparsify.sparsify_errors;
    config.sparsify_base_degree = config_sparsify.sparsify_base_degree;
    config.sparsify_max_degree = config_sparsify.sparsify_max_degree;
    config.sparsify_reactivate_limit = config_sparsify.sparsify_reactivate_limit;

    // One GPU: one decoder.  --gpus N: a tsr_group — one decoder, host thread and NCCL communicator per device, batches
    // dealt out in interleaved chunks, counters summed by one all-reduce (the reference's analogue is one decoder per
    // host thread, tesseract_main.cc:559-571).
    std::unique_ptr<TesseractDecoder> single;
    std::unique_ptr<DecoderGroup> group;
    if (a.gpus == 1) single = std::make_unique<TesseractDecoder>(config);
    else group = std::make_unique<DecoderGroup>(config, a.gpus);
    tsr_decoder* first = single ? single->handle() : group->member(0);
    const uint64_t D = tsr_num_detectors(first), O = tsr_num_observables(first);
    const uint64_t nb = std::max<uint64_t>((D + 7) / 8, 1), ob = (O + 7) / 8;

    // Shots.
    uint64_t shots = 0;
    std::vector<uint8_t> dets, obs_true;
    bool has_obs = false;
    if (a.sample_num_shots) {
      shots = a.sample_num_shots;
      dets.assign((size_t)(shots * nb), 0);
      obs_true.assign((size_t)(shots * ob) + 1, 0);
      check(tsr_sample_dem_b8(dem_text.c_str(), a.sample_seed, shots, dets.data(), obs_true.data()));
      has_obs = true;
    } else {
      if (a.in_includes_appended_observables) {
        uint64_t n = 0;
        std::vector<uint8_t> both = read_shots(a.in_fname, a.in_format, D + O, n, D);
        const uint64_t bb = (D + O + 7) / 8;
        shots = n;
        dets.assign((size_t)(shots * nb), 0);
        obs_true.assign((size_t)(shots * ob) + 1, 0);
        for (uint64_t s = 0; s < shots; ++s)
          for (uint64_t i = 0; i < D + O; ++i)
            if ((both[s * bb + i / 8] >> (i % 8)) & 1) {
              if (i < D) dets[s * nb + i / 8] |= (uint8_t)(1u << (i % 8));
              else obs_true[s * ob + (i - D) / 8] |= (uint8_t)(1u << ((i - D) % 8));
            }
        has_obs = true;
      } else {
        dets = read_shots(a.in_fname, a.in_format, D, shots);
        if (!a.obs_in_fname.empty()) {
          uint64_t n = 0;
          obs_true = read_shots(a.obs_in_fname, a.obs_in_format, O, n);
          if (n != shots) throw std::invalid_argument("--obs_in has a different number of shots than --in");
          obs_true.push_back(0);
          has_obs = true;
        }
      }
    }
    // --shot-range-begin/--shot-range-end (tesseract_main.cc:307-316)
    if (a.shot_range_begin || a.shot_range_end) {
      if (a.shot_range_end < a.shot_range_begin || a.shot_range_end > shots) throw std::invalid_argument("Invalid shot range");
      auto keep = [](std::vector<uint8_t>& v, uint64_t b, uint64_t e) {  // v = v[b:e), without aliasing assign()
        v.erase(v.begin() + (std::ptrdiff_t)e, v.end());
        v.erase(v.begin(), v.begin() + (std::ptrdiff_t)b);
      };
      keep(dets, a.shot_range_begin * nb, a.shot_range_end * nb);
      if (has_obs) keep(obs_true, a.shot_range_begin * ob, a.shot_range_end * ob);
      obs_true.push_back(0);
      shots = a.shot_range_end - a.shot_range_begin;
    }

    std::vector<uint8_t> obs_pred((size_t)(shots * ob) + 1, 0), low(shots, 0);
    std::vector<double> cost(shots, 0);
    std::ofstream out_file;
    if (!a.out_fname.empty() && a.out_fname != "-") {
      out_file.open(a.out_fname, std::ios::binary);
      if (!out_file) throw std::invalid_argument("Failed to open " + a.out_fname);
    }
    std::ostream& out = a.out_fname == "-" ? std::cout : static_cast<std::ostream&>(out_file);
    ShotWriter out_writer(out, a.out_format, O, 'L');

    uint64_t num_errors = 0, num_low_confidence = 0, done = 0;
    double total_time_seconds = 0;
    const bool want_use = !a.dem_out_fname.empty();
    std::vector<uint64_t> error_use(want_use ? (size_t)tsr_num_dem_errors(first) : 0, 0), scratch_use(error_use.size(), 0);
    std::vector<uint64_t> batch_off, batch_idx;
    uint64_t reduced_shots = 0, reduced_errors = 0, reduced_low = 0;  // the group's all-reduced counters (cross-check)
    // Batches in shot order (so --max-errors stops where the reference's in-order consumer would, up to a batch).
    for (uint64_t at = 0; at < shots && num_errors < a.max_errors; at += a.batch) {
      const uint64_t n = std::min<uint64_t>(a.batch, shots - at);
      const auto t0 = std::chrono::steady_clock::now();
      if (single) {
        single->decode_batch_b8(dets.data() + at * nb, nb, n, obs_pred.data() + at * ob, cost.data() + at, low.data() + at, want_use);
        if (want_use) single->last_batch_errors(n, batch_off, batch_idx);
      } else {
        uint64_t tally[3] = {0, 0, 0};
        group->decode_batch_b8(dets.data() + at * nb, nb, n, has_obs ? obs_true.data() + at * ob : nullptr, obs_pred.data() + at * ob,
                               cost.data() + at, low.data() + at, tally, want_use ? scratch_use.data() : nullptr);
        if (want_use) group->last_batch_errors(n, batch_off, batch_idx);
        reduced_shots += tally[0];
        reduced_errors += tally[1];
        reduced_low += tally[2];
      }
      total_time_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      for (uint64_t s = at; s < at + n; ++s) {  // consume(shot), tesseract_main.cc:592-616
        if (!a.out_fname.empty()) out_writer.write(obs_pred.data() + s * ob);
        const bool match = !has_obs || std::memcmp(obs_pred.data() + s * ob, obs_true.data() + s * ob, (size_t)ob) == 0;
        if (low[s]) ++num_low_confidence;
        else if (!match) ++num_errors;
        if (want_use && match) {  // error_use: errors of the shots whose prediction matches (tesseract_main.cc:586-590)
          const uint64_t k = s - at;
          for (uint64_t q = batch_off[k]; q < batch_off[k + 1]; ++q) ++error_use[batch_idx[q]];
        }
        ++done;
        if (a.print_stats) {
          std::cout << "num_shots = " << done << " num_low_confidence = " << num_low_confidence;
          if (has_obs) std::cout << " num_errors = " << num_errors;
          std::cout << " total_time_seconds = " << total_time_seconds << std::endl;
          std::cout << "cost = " << cost[s] << std::endl;
        }
        if (has_obs && num_errors >= a.max_errors) break;
      }
    }

    // With every batch consumed to its end, the in-order host accounting and the all-reduced device-side counters are
    // the same numbers; a difference means the collective or the scatter lost shots.
    if (group && done == reduced_shots && (num_low_confidence != reduced_low || (has_obs && num_errors != reduced_errors)))
      throw std::runtime_error("multi-GPU counters disagree with the per-shot accounting");

    if (!a.out_fname.empty()) out_writer.finish();
    if (want_use) {  // tesseract_main.cc:625-639
      const uint64_t usage_shots = has_obs ? done - num_errors : done;
      char* text = nullptr;
      check(tsr_dem_from_counts(dem_text.c_str(), error_use.data(), error_use.size(), usage_shots, &text));
      std::ofstream f(a.dem_out_fname);
      if (!f) throw std::invalid_argument("Failed to open " + a.dem_out_fname);
      f << text << '\n';
      tsr_free(text);
    }

    bool print_final_stats = true;
    if (!a.stats_out_fname.empty()) {  // tesseract_main.cc:657-690
      std::ostringstream js;
      auto q = [](const std::string& s) { return "\"" + s + "\""; };
      js << "{\"circuit_path\":" << q(a.circuit_path) << ",\"dem_path\":" << q(a.dem_path) << ",\"max_errors\":" << a.max_errors
         << ",\"sample_seed\":" << a.sample_seed << ",\"det_beam\":" << a.det_beam << ",\"det_penalty\":" << a.det_penalty
         << ",\"beam_climbing\":" << (a.beam_climbing ? "true" : "false") << ",\"no_revisit_dets\":" << (a.no_revisit_dets ? "true" : "false")
         << ",\"at_most_two_errors_per_detector\":" << (a.at_most_two ? "true" : "false") << ",\"pqlimit\":" << a.pqlimit
         << ",\"num_det_orders\":" << a.num_det_orders << ",\"det_order_seed\":" << a.det_order_seed
         << ",\"total_time_seconds\":" << total_time_seconds << ",\"num_errors\":" << (has_obs ? std::to_string(num_errors) : "null")
         << ",\"num_low_confidence\":" << num_low_confidence << ",\"num_shots\":" << done << ",\"num_threads\":" << a.num_threads
         << ",\"num_gpus\":" << a.gpus << ",\"sample_num_shots\":" << a.sample_num_shots
         // tesseract_main.cc:676-679; the limit is the decoder's effective value (:641-655)
         << ",\"sparsify_errors\":" << (config.sparsify_errors ? "true" : "false") << ",\"sparsify_base_degree\":" << config.sparsify_base_degree
         << ",\"sparsify_max_degree\":" << config.sparsify_max_degree
         << ",\"sparsify_reactivate_limit\":" << tsr_sparsify_reactivate_limit(first) << "}";
      if (a.stats_out_fname == "-") {
        std::cout << js.str() << std::endl;
        print_final_stats = false;
      } else {
        std::ofstream f(a.stats_out_fname);
        f << js.str() << std::endl;
      }
    }
    if (print_final_stats) {  // tesseract_main.cc:691-699
      std::cout << "num_shots = " << done << " num_low_confidence = " << num_low_confidence;
      if (has_obs) std::cout << " num_errors = " << num_errors;
      std::cout << " total_time_seconds = " << total_time_seconds << std::endl;
    }
  } catch (const std::exception& e) {
    std::cerr << "tesseract: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
