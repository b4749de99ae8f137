// SPDX-License-Identifier: MIT
//
// verify-batch — the batched counterpart of `simfony run <prog> --witness <file.wit>`
// (simfony-cli/src/main.rs:49-60,163-209).  The reference verifies one proof per process and reports
// accept as exit code 0 / reject as "Error: ..." + exit 1 (main.rs:271-274); this tool verifies any number
// of witnesses of one of the two reference programs on 1..k B200s and prints one accept / reject line per
// witness, exiting 0 iff every witness is accepted.  (north_star asks for this as a Rust sub-command inside
// simfony-cli; there is no Rust toolchain in this environment, so it is a C++ tool over the same C-ABI —
// INTEGRATION.md shows the Rust binding.)
//
//   verify-batch --program {stwo,stark101} [--preset {prod,testing}] [--mode {ref-literal,prover-consistent}]
//                (--witness a.wit [b.wit ...] | --witness-dir DIR) [--replicate N] [--gpus K] [--trace out.json] [--cost] [--quiet] [--host-pack]
//                [--queries Q]
//
// --cost (stwo): after verifying, prints what the reference PROGRAM executes per proof — sha_256_ctx_8_* jet calls and compressions, M31
// multiplications / additions / inversions, eq_256 — from the cost model of include/ssym.h (ssym_stwo_cost) fed with the queries each
// transcript drew; the dynamic counterpart of the `Node bounds` line `simfony run` prints (simfony-cli/src/main.rs:193-203).
//
// Witnesses are tokenised and packed ON THE GPU (ssym_stwo_verify_wit_batch / ssym_stark101_verify_wit_batch: the `.wit` text is what crosses
// PCIe); --host-pack (and --trace, which needs the packed records on the host) uses the host parser + the packed-batch entry points instead.
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ssym.h"

static bool read_file(const std::string &path, std::string &out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

static void usage() {
    fprintf(stderr,
            "usage: verify-batch --program {stwo,stark101} [--preset {prod,testing}] [--columns {4,8,16}] [--mode {ref-literal,prover-consistent}]\n"
            "                    (--witness a.wit [b.wit ...] | --witness-dir DIR) [--replicate N] [--gpus K] [--trace out.json] [--cost] [--dedup-queries] [--quiet] [--host-pack]\n"
            "                    [--queries Q]   (stark101: every Q consecutive witnesses are ONE proof, witness k under the (k+1)-th query draw)\n");
}

static std::string hex_digest(const uint32_t *w) {
    char buf[65];
    for (int i = 0; i < 8; i++) snprintf(buf + 8 * i, 9, "%08x", w[i]);
    return std::string(buf, 64);
}
static std::string arr(const uint32_t *w, int n) {
    std::string s = "[";
    for (int i = 0; i < n; i++) s += (i ? "," : "") + std::to_string(w[i]);
    return s + "]";
}

int main(int argc, char **argv) {
    std::string program, preset = "prod", mode = "ref-literal", witness_dir, trace_path;
    std::vector<std::string> witnesses;
    size_t replicate = 1;
    int gpus = 1;
    uint32_t columns = SSYM_NUM_COLUMNS; // NUM_COLUMNS of the program the witnesses were made for (config.simf:14)
    bool quiet = false, host_pack = false, want_cost = false, dedup = false;
    uint32_t queries = 1; // stark101 multi-query (include/ssym.h ssym_stark101_verify_multi_batch): witnesses per proof
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&](const char *what) -> std::string {
            if (i + 1 >= argc) { fprintf(stderr, "Error: %s needs a value\n", what); exit(2); }
            return argv[++i];
        };
        if (a == "--program") program = next("--program");
        else if (a == "--preset") preset = next("--preset");
        else if (a == "--mode") mode = next("--mode");
        else if (a == "--witness") {
            witnesses.push_back(next("--witness"));
            while (i + 1 < argc && argv[i + 1][0] != '-') witnesses.push_back(argv[++i]);
        } else if (a == "--witness-dir") witness_dir = next("--witness-dir");
        else if (a == "--replicate") replicate = strtoull(next("--replicate").c_str(), nullptr, 10);
        else if (a == "--gpus") gpus = atoi(next("--gpus").c_str());
        else if (a == "--columns") columns = (uint32_t)strtoul(next("--columns").c_str(), nullptr, 10);
        else if (a == "--trace") trace_path = next("--trace");
        else if (a == "--quiet") quiet = true;
        else if (a == "--cost") want_cost = true;
        else if (a == "--dedup-queries") dedup = true;
        else if (a == "--host-pack") host_pack = true;
        else if (a == "--queries") queries = (uint32_t)strtoul(next("--queries").c_str(), nullptr, 10);
        else if (a == "--help" || a == "-h") { usage(); return 0; }
        else { fprintf(stderr, "Error: unknown argument %s\n", a.c_str()); usage(); return 2; }
    }
    if (!witness_dir.empty()) {
        DIR *d = opendir(witness_dir.c_str());
        if (!d) { fprintf(stderr, "Error: Failed to read witness dir: %s\n", witness_dir.c_str()); return 1; }
        while (dirent *e = readdir(d)) {
            std::string nm = e->d_name;
            if (nm.size() > 4 && nm.substr(nm.size() - 4) == ".wit") witnesses.push_back(witness_dir + "/" + nm);
        }
        closedir(d);
        std::sort(witnesses.begin(), witnesses.end());
    }
    if ((program != "stwo" && program != "stark101") || witnesses.empty() || replicate < 1 || gpus < 1) { usage(); return 2; }
    if (queries != 1) {
        if (program != "stark101" || queries < 1 || queries > SSYM_S101_MAX_ORDINAL + 1u || witnesses.size() % queries) {
            fprintf(stderr, "Error: --queries Q groups the witnesses of --program stark101 by Q (1 .. 256); %zu witnesses given\n", witnesses.size());
            return 2;
        }
        host_pack = true; // the query ordinal is a word of the packed record
    }
    uint32_t mode_id;
    if (mode == "ref-literal") mode_id = SSYM_MODE_REF_LITERAL;
    else if (mode == "prover-consistent") mode_id = SSYM_MODE_PROVER_CONSISTENT;
    else { usage(); return 2; }
    if (dedup) mode_id |= SSYM_MODE_QUERY_DEDUP; // fri/queries.simf:41: sort the queries and drop duplicates (include/ssym.h)

    const size_t n_files = witnesses.size(), n = n_files * replicate;
    std::vector<uint8_t> parse_reject(n_files, 0); // ill-typed or ill-shaped witness: simfony would refuse it -> reject
    std::vector<uint32_t> accept((n + 31) / 32, 0), status(n, 0);
    if (want_cost && program != "stwo") { fprintf(stderr, "Error: --cost models the stwo program only\n"); return 2; }
    const bool want_trace = !trace_path.empty() || want_cost; // the cost model needs the queries each transcript drew
    auto t0 = std::chrono::steady_clock::now();

    ssym_stwo_config_t cfg{};
    ssym_stwo_layout_t lo{};
    std::vector<uint32_t> packed;          // stwo
    std::vector<uint32_t> blob;            // stark101
    std::vector<uint64_t> offsets;
    std::vector<ssym_stwo_trace_t> traces;
    std::vector<ssym_s101_trace_t> traces101;
    const bool gpu_ingest = !want_trace && !host_pack;
    char *wit_text = nullptr;           // GPU ingestion: the concatenated witness texts, read straight into page-locked memory
    std::vector<uint64_t> wit_offsets;
    std::vector<uint32_t> wit_flags;
    if (gpu_ingest) {
        if (program == "stwo" && (ssym_stwo_config_preset(preset.c_str(), mode_id, &cfg) || (cfg.n_columns = columns, ssym_stwo_layout(&cfg, &lo)))) { fprintf(stderr, "Error: %s\n", ssym_last_error()); return 2; }
        wit_offsets.push_back(0);
        for (size_t f = 0; f < n_files; f++) {
            struct stat st;
            if (stat(witnesses[f].c_str(), &st) != 0 || !S_ISREG(st.st_mode)) { fprintf(stderr, "Error: Failed to read witness file: %s\n", witnesses[f].c_str()); return 1; }
            wit_offsets.push_back(wit_offsets.back() + (uint64_t)st.st_size);
        }
        const size_t one = wit_offsets.back();
        wit_text = static_cast<char *>(ssym_pinned_alloc(one * replicate + 16));
        bool pinned = wit_text != nullptr;
        if (!pinned) wit_text = static_cast<char *>(malloc(one * replicate + 16)); // no GPU / no pinned memory: ssym_create reports it below
        if (!wit_text) { fprintf(stderr, "Error: out of memory\n"); return 1; }
        for (size_t f = 0; f < n_files; f++) {
            FILE *fp = fopen(witnesses[f].c_str(), "rb");
            const size_t want = (size_t)(wit_offsets[f + 1] - wit_offsets[f]);
            if (!fp || fread(wit_text + wit_offsets[f], 1, want, fp) != want) { fprintf(stderr, "Error: Failed to read witness file: %s\n", witnesses[f].c_str()); return 1; }
            fclose(fp);
        }
        for (size_t r = 1; r < replicate; r++) {
            memcpy(wit_text + r * one, wit_text, one);
            for (size_t f = 0; f < n_files; f++) wit_offsets.push_back(r * one + wit_offsets[f + 1]);
        }
        wit_flags.assign(n, 0);
    } else if (program == "stwo") {
        if (ssym_stwo_config_preset(preset.c_str(), mode_id, &cfg) || (cfg.n_columns = columns, ssym_stwo_layout(&cfg, &lo))) { fprintf(stderr, "Error: %s\n", ssym_last_error()); return 2; }
        packed.assign(n * (size_t)lo.stride_words, 0);
        for (size_t f = 0; f < n_files; f++) {
            std::string text;
            if (!read_file(witnesses[f], text)) { fprintf(stderr, "Error: Failed to read witness file: %s\n", witnesses[f].c_str()); return 1; }
            int shape = 0;
            int rc = ssym_stwo_pack_wit(&cfg, text.data(), text.size(), packed.data() + f * (size_t)lo.stride_words, &shape);
            parse_reject[f] = (rc != SSYM_OK || shape) ? 1 : 0;
        }
        for (size_t r = 1; r < replicate; r++) memcpy(packed.data() + r * n_files * (size_t)lo.stride_words, packed.data(), n_files * (size_t)lo.stride_words * 4);
        if (want_trace) traces.resize(n);
    } else {
        offsets.push_back(0);
        for (size_t f = 0; f < n_files; f++) {
            std::string text;
            if (!read_file(witnesses[f], text)) { fprintf(stderr, "Error: Failed to read witness file: %s\n", witnesses[f].c_str()); return 1; }
            std::vector<uint32_t> rec(20 + 8 * 3 * 31 + 31 * (16 + 8 * 62));
            size_t words = rec.size();
            int rc = ssym_s101_pack_wit(text.data(), text.size(), rec.data(), &words);
            if (rc != SSYM_OK) { // keep a minimal malformed record so indices stay aligned
                parse_reject[f] = 1;
                words = 20;
                std::fill(rec.begin(), rec.begin() + 20, 0u);
                rec[0] = 20;
            }
            if (rc == SSYM_OK) rec[6] = (uint32_t)(f % queries); // query ordinal: witness k of a proof is verified under the (k+1)-th draw
            blob.insert(blob.end(), rec.begin(), rec.begin() + words);
            offsets.push_back(blob.size());
        }
        const size_t one = blob.size();
        for (size_t r = 1; r < replicate; r++) {
            blob.insert(blob.end(), blob.begin(), blob.begin() + one);
            for (size_t f = 0; f < n_files; f++) offsets.push_back(r * one + offsets[f + 1]);
        }
        if (want_trace) traces101.resize(n);
    }
    auto t1 = std::chrono::steady_clock::now();

    // contiguous shards over the GPUs, one host thread and one handle per GPU; shard sizes are multiples of 32
    const size_t unit = 32 * (size_t)queries; // shard sizes are multiples of 32 proofs
    gpus = (int)std::min<size_t>((size_t)gpus, (n + unit - 1) / unit);
    std::vector<int> rcs(gpus, 0);
    std::vector<std::string> errs(gpus);
    std::vector<std::thread> threads;
    const size_t per = (((n + gpus - 1) / gpus) + unit - 1) / unit * unit;
    std::vector<uint32_t> proof_accept((n / queries + 31) / 32, 0); // --queries: one bit per proof
    for (int g = 0; g < gpus; g++) {
        threads.emplace_back([&, g]() {
            const size_t b = std::min(n, g * per), e = std::min(n, b + per);
            if (b >= e) return;
            ssym_ctx_t *ctx = nullptr;
            int rc = ssym_create(g, &ctx);
            if (rc == SSYM_OK) {
                if (gpu_ingest) {
                    std::vector<uint64_t> offs(wit_offsets.begin() + b, wit_offsets.begin() + e + 1);
                    const uint64_t base = offs[0];
                    for (auto &o : offs) o -= base;
                    rc = program == "stwo" ? ssym_stwo_verify_wit_batch(ctx, &cfg, wit_text + base, offs.data(), e - b, accept.data() + b / 32,
                                                                        status.data() + b, wit_flags.data() + b, SSYM_MEM_HOST)
                                           : ssym_stark101_verify_wit_batch(ctx, wit_text + base, offs.data(), e - b, accept.data() + b / 32,
                                                                            status.data() + b, wit_flags.data() + b, SSYM_MEM_HOST);
                } else if (program == "stwo")
                    rc = ssym_stwo_verify_batch(ctx, &cfg, packed.data() + b * (size_t)lo.stride_words, e - b, accept.data() + b / 32, status.data() + b,
                                                want_trace ? traces.data() + b : nullptr, SSYM_MEM_HOST);
                else {
                    std::vector<uint64_t> offs(offsets.begin() + b, offsets.begin() + e + 1);
                    const uint64_t base = offs[0];
                    for (auto &o : offs) o -= base;
                    if (queries > 1)
                        rc = ssym_stark101_verify_multi_batch(ctx, blob.data() + base, offs.data(), (e - b) / queries, queries, proof_accept.data() + b / queries / 32,
                                                              status.data() + b, want_trace ? traces101.data() + b : nullptr, SSYM_MEM_HOST);
                    else
                        rc = ssym_stark101_verify_batch(ctx, blob.data() + base, offs.data(), e - b, accept.data() + b / 32, status.data() + b,
                                                        want_trace ? traces101.data() + b : nullptr, SSYM_MEM_HOST);
                }
            }
            if (rc != SSYM_OK) errs[g] = ssym_last_error();
            rcs[g] = rc;
            ssym_destroy(ctx);
        });
    }
    for (auto &t : threads) t.join();
    for (int g = 0; g < gpus; g++)
        if (rcs[g] != SSYM_OK) { fprintf(stderr, "Error: GPU %d: %s\n", g, errs[g].c_str()); return 3; }
    auto t2 = std::chrono::steady_clock::now();

    size_t n_accept = 0;
    for (size_t i = 0; i < n; i++) {
        const size_t f = i % n_files;
        if (gpu_ingest && wit_flags[i] != SSYM_WIT_OK) parse_reject[f] = 1;
        bool ok = (queries > 1 ? status[i] == 0 && ((proof_accept[i / queries / 32] >> (i / queries % 32)) & 1) : ((accept[i / 32] >> (i % 32)) & 1)) && !parse_reject[f];
        if (queries > 1 && !ok && status[i] == 0 && !parse_reject[f]) status[i] |= SSYM_S101_ST_GROUP; // rejected with its proof: another of its records failed
        if (parse_reject[f]) status[i] |= SSYM_ST_SHAPE;
        n_accept += ok;
        if (!quiet) printf("%s %s%s\n", ok ? "accept" : "reject", witnesses[f].c_str(), ok ? "" : (" status=0x" + [&] { char b[16]; snprintf(b, sizeof b, "%08x", status[i]); return std::string(b); }()).c_str());
    }
    const double pack_s = std::chrono::duration<double>(t1 - t0).count(), gpu_s = std::chrono::duration<double>(t2 - t1).count();
    fprintf(stderr, "verify-batch: %zu proofs, %zu accepted, %zu rejected; %s %.3f s, %s (%d GPU%s, host buffers) %.3f s\n", n, n_accept,
            n - n_accept, gpu_ingest ? "read" : "parse+pack", pack_s, gpu_ingest ? "tokenise+pack+verify" : "verify", gpus, gpus > 1 ? "s" : "", gpu_s);

    if (want_cost) {
        static const char *const names[SSYM_COST_FIELDS] = {"sha_compressions", "sha_init", "sha_add_4", "sha_add_8", "sha_add_32", "sha_finalize", "sha_bytes",
                                                            "m31_mul", "m31_add", "m31_neg", "m31_inv", "eq_256", "point_from_index", "draw_retries"};
        uint64_t total[SSYM_COST_FIELDS] = {0};
        for (size_t i = 0; i < n; i++) {
            ssym_cost_t c;
            if (parse_reject[i % n_files] || ssym_stwo_cost(&cfg, traces[i].queries, traces[i].n_queries_used, traces[i].draw_retries, &c) != SSYM_OK) continue;
            const uint64_t *v = reinterpret_cast<const uint64_t *>(&c);
            if (!quiet) printf("cost %s", witnesses[i % n_files].c_str());
            for (int k = 0; k < SSYM_COST_FIELDS; k++) {
                total[k] += v[k];
                if (!quiet) printf(" %s=%llu", names[k], (unsigned long long)v[k]);
            }
            if (!quiet) printf("\n");
        }
        fprintf(stderr, "verify-batch: program cost over %zu proofs (what verify_proof executes, run to its end):", n);
        for (int k = 0; k < SSYM_COST_FIELDS; k++) fprintf(stderr, " %s=%llu", names[k], (unsigned long long)total[k]);
        // the jets underneath the field functions (fields/m31.simf:22-45,117-132)
        fprintf(stderr, "\nverify-batch: as jets: multiply_32=%llu modulo_64=%llu add_32=%llu modulo_32=%llu subtract_32=%llu is_zero_32=%llu eq_256=%llu "
                        "sha_256_ctx_8_{init=%llu,add_4=%llu,add_8=%llu,add_32=%llu,finalize=%llu}\n",
                (unsigned long long)total[7], (unsigned long long)total[7], (unsigned long long)total[8], (unsigned long long)total[8], (unsigned long long)total[9],
                (unsigned long long)total[10], (unsigned long long)total[11], (unsigned long long)total[1], (unsigned long long)total[2], (unsigned long long)total[3],
                (unsigned long long)total[4], (unsigned long long)total[5]);
    }
    if (!trace_path.empty()) {
        std::ofstream o(trace_path);
        o << "[\n";
        for (size_t i = 0; i < n; i++) {
            if (program == "stwo") {
                const ssym_stwo_trace_t &t = traces[i];
                o << " {\"witness\": \"" << witnesses[i % n_files] << "\", \"status\": " << status[i] << ", \"first_fail\": " << t.first_fail
                  << ", \"digest_commit\": \"" << hex_digest(t.digest_commit) << "\", \"cp_alpha\": " << arr(t.cp_alpha, 4)
                  << ", \"oods_x\": " << arr(t.oods_x, 4) << ", \"oods_y\": " << arr(t.oods_y, 4) << ", \"cp_eval\": " << arr(t.cp_eval, 4)
                  << ", \"cp_sampled\": " << arr(t.cp_sampled, 4) << ", \"digest_oods\": \"" << hex_digest(t.digest_oods) << "\", \"deep_alpha\": "
                  << arr(t.deep_alpha, 4) << ", \"digest_fri\": \"" << hex_digest(t.digest_fri) << "\", \"digest_pow\": \"" << hex_digest(t.digest_pow)
                  << "\", \"queries\": " << arr(t.queries, (int)cfg.n_queries) << ", \"fri_answer\": [";
                for (uint32_t q = 0; q < cfg.n_queries; q++) o << (q ? "," : "") << arr(t.fri_answer[q], 4);
                o << "], \"folded_last\": [";
                for (uint32_t q = 0; q < cfg.n_queries; q++) o << (q ? "," : "") << arr(t.folded[cfg.n_fri_layers][q], 4);
                o << "]}";
            } else {
                const ssym_s101_trace_t &t = traces101[i];
                o << " {\"witness\": \"" << witnesses[i % n_files] << "\", \"status\": " << status[i] << ", \"alpha\": " << arr(t.alpha, 3)
                  << ", \"idx\": " << t.idx << ", \"x\": " << t.x << ", \"cp0\": " << t.cp0 << ", \"n_layers\": " << t.n_layers << ", \"cp_ev\": "
                  << arr(t.cp_ev, (int)std::min<uint32_t>(t.n_layers, 31) + 1) << "}";
            }
            o << (i + 1 < n ? ",\n" : "\n");
        }
        o << "]\n";
    }
    if (n_accept != n) {
        fprintf(stderr, "Error: Failed to run program: %zu of %zu witnesses rejected\n", n - n_accept, n);
        return 1;
    }
    return 0;
}
