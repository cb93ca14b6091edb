// lsbsort -- command-line driver with the reference's knobs and report lines.
//
// Mirrors main() of the reference (mpi/mpi_lsbsort.cpp:587-743): --n, --print, --verify,
// --no-verify (:594-611), the same stdout lines (:619-620,:646,:662,:684,:697-699,:712), the
// same default problem size (:593) and verify default (n < 128Mi, :609-611).  Where the
// reference is launched as `mpirun -n R ./mpi_lsbsort`, this driver takes --gpus G (one
// worker process per GPU, forked here) and --ranks R (number of pcg64 streams the generator
// uses, i.e. the R of the run being reproduced; default G).  All compute goes through the C
// ABI in include/lsbsort.h; nothing here touches CUDA directly.
//
// Verification differs in mechanism, not in strength: instead of gathering everything on
// rank 0 and comparing with std::stable_sort (:710-739, O(n) host memory), the shards are
// checked on the GPUs for strictly increasing (key,val) within and across shards plus an
// unchanged multiset hash, which is equivalent because val is the unique input index.
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lsbsort.h"

namespace {

struct Options {
  int64_t n = 100 * 1000 * 1000;  // :593
  bool print = false, verify = false, verify_set = false;
  int gpus = 1, ranks = 0, radix = 16, and_draws = 1;
  bool one_pass = false;
  uint64_t key_mask = ~0ULL, seed_base = 0;
};

void die(const char* what, lsb_ctx* c, int rc) {
  std::fprintf(stderr, "lsbsort: %s failed: %s (%s)\n", what, lsb_status_string(rc), lsb_last_error(c));
  std::exit(2);
}

bool read_all(int fd, void* buf, size_t n) {
  char* p = static_cast<char*>(buf);
  while (n) {
    ssize_t r = read(fd, p, n);
    if (r <= 0) return false;
    p += r;
    n -= (size_t)r;
  }
  return true;
}
bool write_all(int fd, const void* buf, size_t n) {
  const char* p = static_cast<const char*>(buf);
  while (n) {
    ssize_t r = write(fd, p, n);
    if (r <= 0) return false;
    p += r;
    n -= (size_t)r;
  }
  return true;
}
void send_text(int fd, const std::string& s) {
  uint64_t len = s.size();
  write_all(fd, &len, sizeof(len));
  write_all(fd, s.data(), s.size());
}
std::string recv_text(int fd) {
  uint64_t len = 0;
  if (!read_all(fd, &len, sizeof(len))) return std::string();
  std::string s(len, '\0');
  read_all(fd, &s[0], len);
  return s;
}

// DistributedArray::print (:171-200): first 10 elements of this shard, "A[g] = (key hex,val)"
std::string shard_lines(lsb_ctx* c, int64_t per_print) {
  int64_t per = 0, here = 0, first = 0;
  lsb_shard_info(c, &per, &here, &first);
  const int64_t k = here < per_print ? here : per_print;
  std::vector<lsb_elt> e((size_t)(k > 0 ? k : 1));
  int rc = lsb_download(c, e.data(), 0, k);
  if (rc) die("lsb_download", c, rc);
  std::string out;
  char line[128];
  for (int64_t i = 0; i < k; i++) {
    std::snprintf(line, sizeof(line), "A[%" PRId64 "] = (%016" PRIx64 ",%" PRIu64 ")\n", first + i, e[i].key, e[i].val);
    out += line;
  }
  if (k < here) out += "...\n";
  return out;
}

// one worker == one rank of the reference; `up`/`down` talk to the coordinating parent
int worker(const Options& o, int g, int up, int down) {
  lsb_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.n = o.n;
  cfg.ranks = o.ranks ? o.ranks : o.gpus;
  cfg.world_size = o.gpus;
  cfg.world_rank = g;
  cfg.device = g;
  cfg.radix_bits = o.radix;
  cfg.and_draws = o.and_draws;
  cfg.seed_base = o.seed_base;
  cfg.key_mask = o.key_mask;
  cfg.flags = LSB_FLAG_PHASE_EVENTS | (o.one_pass ? LSB_FLAG_ONE_PASS : 0u);
  lsb_ctx* c = nullptr;
  int rc = lsb_create(&c, &cfg);
  if (rc) die("lsb_create", nullptr, rc);
  if (o.gpus > 1) {
    char id[LSB_COMM_ID_BYTES];
    if (g == 0) {
      if ((rc = lsb_comm_unique_id(id))) die("lsb_comm_unique_id", nullptr, rc);
      write_all(up, id, sizeof(id));
    }
    if (!read_all(down, id, sizeof(id))) return 3;
    if ((rc = lsb_comm_init(c, id))) die("lsb_comm_init", c, rc);
  }
  const bool root = (g == 0);
  auto say = [&](const std::string& s) { if (root) send_text(up, s); };
  const auto npos_line = [&](const char* fmt, double v) { char b[160]; std::snprintf(b, sizeof(b), fmt, v); return std::string(b); };

  say("Generating random values\n");
  auto t0 = std::chrono::steady_clock::now();
  if ((rc = lsb_generate(c))) die("lsb_generate", c, rc);
  if ((rc = lsb_barrier(c))) die("lsb_barrier", c, rc);
  std::chrono::duration<double> el = std::chrono::steady_clock::now() - t0;
  say(npos_line("Generated random values in %g s\n", el.count()));

  if (o.print) send_text(up, "\x01" + shard_lines(c, 10));
  uint64_t before[4] = {0, 0, 0, 0};
  lsb_verify vin;
  if (o.verify) {  // the reference saves a copy of the input here (:673-679); a multiset hash suffices
    rc = lsb_verify_device(c, &vin);
    if (rc && rc != LSB_ERR_VERIFY) die("lsb_verify_device", c, rc);
    std::memcpy(before, vin.checksum, sizeof(before));
  }

  say("Sorting\n");
  if ((rc = lsb_barrier(c))) die("lsb_barrier", c, rc);
  t0 = std::chrono::steady_clock::now();
  lsb_stats st;
  if ((rc = lsb_sort(c, &st))) die("lsb_sort", c, rc);
  if ((rc = lsb_barrier(c))) die("lsb_barrier", c, rc);
  el = std::chrono::steady_clock::now() - t0;
  if (root) {
    char b[512];
    std::snprintf(b, sizeof(b), "Sorted %" PRId64 " values in %g\nThat's %g M elements sorted / s\n", o.n, el.count(),
                  o.n / el.count() / 1000.0 / 1000.0);
    say(b);
    const double per_launch = st.partition_launches ? st.partition_ms / st.partition_launches : 0.0;
    std::snprintf(b, sizeof(b),
                  "[b200] device time %.3f ms: count %.3f ms, scan+collectives %.3f ms, %d partition launches "
                  "%.3f ms each = %.0f GB/s per launch (32 B/element); %d passes x %d-bit digits on %d GPU(s)\n",
                  st.device_ms, st.hist_ms, st.scan_ms, (int)st.partition_launches, per_launch,
                  per_launch > 0 ? st.elements * 32.0 / (per_launch * 1e-3) / 1e9 : 0.0, st.passes, o.radix, o.gpus);
    say(b);
  }
  if (o.print) send_text(up, "\x01" + shard_lines(c, 10));
  int status = 0;
  if (o.verify) {
    say("Verifying\n");
    lsb_verify v;
    rc = lsb_verify_device(c, &v);
    if (rc && rc != LSB_ERR_VERIFY) die("lsb_verify_device", c, rc);
    const bool same = std::memcmp(before, v.checksum, sizeof(before)) == 0 && v.elements == o.n;
    if (rc == LSB_ERR_VERIFY || !same) {
      status = 1;
      char b[256];
      std::snprintf(b, sizeof(b), "Verification FAILED: %" PRId64 " order violations, multiset %s\n", v.order_violations,
                    same ? "unchanged" : "CHANGED");
      say(b);
    }
  }
  lsb_destroy(c);
  if (root) send_text(up, "\x02");
  return status;
}

}  // namespace

int main(int argc, char** argv) {
  Options o;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { if (i + 1 >= argc) { std::fprintf(stderr, "lsbsort: %s needs a value\n", a.c_str()); std::exit(2); } return argv[++i]; };
    if (a == "--n") o.n = std::stoll(next());
    else if (a == "--print") o.print = true;
    else if (a == "--verify") { o.verify = true; o.verify_set = true; }
    else if (a == "--no-verify") { o.verify = false; o.verify_set = true; }
    else if (a == "--gpus") o.gpus = std::atoi(next());
    else if (a == "--ranks") o.ranks = std::atoi(next());
    else if (a == "--radix") o.radix = std::atoi(next());
    else if (a == "--key-mask") o.key_mask = std::strtoull(next(), nullptr, 0);
    else if (a == "--and-draws") o.and_draws = std::atoi(next());
    else if (a == "--one-pass") o.one_pass = true;
    else if (a == "--seed-base") o.seed_base = std::strtoull(next(), nullptr, 0);
    else if (a == "--help" || a == "-h") {
      std::printf("usage: lsbsort [--n N] [--gpus G] [--ranks R] [--radix BITS] [--verify|--no-verify] [--print]\n"
                  "               [--key-mask M] [--and-draws K] [--seed-base S] [--one-pass]\n");
      return 0;
    }
  }
  if (!o.verify_set) o.verify = (o.n < 128LL * 1024 * 1024);  // :609-611
  if (o.gpus < 1 || o.gpus > LSB_MAX_GPUS) { std::fprintf(stderr, "lsbsort: --gpus must be 1..%d\n", LSB_MAX_GPUS); return 2; }

  // rank 0 of the reference prints these before anything else (:618-622)
  std::printf("Total number of MPI ranks: %d\nProblem size: %" PRId64 "\n", o.ranks ? o.ranks : o.gpus, o.n);
  std::fflush(stdout);

  std::vector<int> up(o.gpus), down(o.gpus);
  std::vector<pid_t> pid(o.gpus);
  for (int g = 0; g < o.gpus; g++) {
    int pu[2], pd[2];
    if (pipe(pu) || pipe(pd)) { std::perror("pipe"); return 2; }
    pid[g] = fork();
    if (pid[g] < 0) { std::perror("fork"); return 2; }
    if (pid[g] == 0) {
      close(pu[0]);
      close(pd[1]);
      for (int h = 0; h < g; h++) { close(up[h]); close(down[h]); }
      _exit(worker(o, g, pu[1], pd[0]));
    }
    close(pu[1]);
    close(pd[0]);
    up[g] = pu[0];
    down[g] = pd[1];
  }
  if (o.gpus > 1) {  // forward the communicator id from worker 0 to everyone
    char id[LSB_COMM_ID_BYTES];
    if (!read_all(up[0], id, sizeof(id))) return 2;
    for (int g = 0; g < o.gpus; g++) write_all(down[g], id, sizeof(id));
  }
  // relay: text from worker 0 verbatim; a \x01 block means "every worker now sends its print
  // block", shown in rank order like the reference's barrier loop (:187-199)
  while (true) {
    std::string s = recv_text(up[0]);
    if (s.empty() || s[0] == '\x02') break;
    if (s[0] == '\x01') {
      if (10LL * o.gpus >= o.n) std::printf("A: displaying all %" PRId64 " elements\n", o.n);
      else std::printf("A: displaying first 10 elements on each rank out of %" PRId64 " elements\n", o.n);
      std::fputs(s.c_str() + 1, stdout);
      for (int g = 1; g < o.gpus; g++) {
        std::string t = recv_text(up[g]);
        if (!t.empty()) std::fputs(t.c_str() + 1, stdout);
      }
    } else {
      std::fputs(s.c_str(), stdout);
    }
    std::fflush(stdout);
  }
  int status = 0;
  for (int g = 0; g < o.gpus; g++) {
    int ws = 0;
    waitpid(pid[g], &ws, 0);
    if (!WIFEXITED(ws) || WEXITSTATUS(ws)) status = WIFEXITED(ws) ? WEXITSTATUS(ws) : 2;
  }
  return status;
}
