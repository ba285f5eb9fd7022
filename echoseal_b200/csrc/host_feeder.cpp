// echoseal_b200/csrc/host_feeder.cpp — native, multi-threaded HOST side of the path.
//
// BASELINE.json north_star keeps four things on the host: key handling (HKDF), the HMAC hop schedule,
// the AES-CTR PN chip generator and the ChaCha20-Poly1305 tag check.  In the reference they are Python
// calls made one frame / one counter at a time (rtwm/crypto.py:14-48, rtwm/utils.py:27-36, 83-132,
// rtwm/detector.py:117-151, 168-233, rtwm/embedder.py:153-168); at GPU speed that is the Amdahl term
// (SURVEY.md section 8f-1), so here they are batched over whole sub-batches of clips / frames and spread
// over the host cores.  Nothing in this file touches the GPU: it produces kernel inputs (candidate
// lists, PN bits, sealed payloads, hop bands) and consumes kernel outputs (CRC-passing candidates).
#include <openssl/evp.h>
#include <openssl/hmac.h>
#include <openssl/sha.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <atomic>
#include <thread>
#include <functional>
#include <mutex>
#include <condition_variable>
#include <vector>
#include <unistd.h>
#include <algorithm>

namespace {

constexpr int FRAME_LEN = 1215, PEAK_LIMIT = 25, MAX_TRIES = 400, TIGHT = 3, WIDE = 200;
constexpr int PN_BYTES = 152, PN_BLOCKS = 10;

// ---------------------------------------------------------------- BLAKE2s (RFC 7693), unkeyed, personalised
static const uint32_t B2S_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                   0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

// digest of a <= 64-byte message, outlen <= 32, 8-byte personalisation
static void blake2s_short(const uint8_t* msg, size_t len, const uint8_t person[8], uint8_t* out, int outlen)
{
    uint32_t h[8], m[16], v[16];
    uint8_t block[64] = {0};
    memcpy(block, msg, len);
    for (int i = 0; i < 8; ++i) h[i] = B2S_IV[i];
    h[0] ^= 0x01010000u ^ (uint32_t)outlen;
    uint32_t p6, p7;
    memcpy(&p6, person, 4); memcpy(&p7, person + 4, 4);
    h[6] ^= p6; h[7] ^= p7;
    for (int i = 0; i < 16; ++i) memcpy(&m[i], block + 4 * i, 4);
    for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = B2S_IV[i]; }
    v[12] ^= (uint32_t)len;      // t0
    v[14] ^= 0xFFFFFFFFu;        // last block
#define B2G(a, b, c, d, x, y) \
    v[a] = v[a] + v[b] + (x); v[d] = rotr32(v[d] ^ v[a], 16); v[c] = v[c] + v[d]; v[b] = rotr32(v[b] ^ v[c], 12); \
    v[a] = v[a] + v[b] + (y); v[d] = rotr32(v[d] ^ v[a], 8);  v[c] = v[c] + v[d]; v[b] = rotr32(v[b] ^ v[c], 7);
    for (int r = 0; r < 10; ++r) {
        const uint8_t* s = B2S_SIGMA[r];
        B2G(0, 4, 8, 12, m[s[0]], m[s[1]]) B2G(1, 5, 9, 13, m[s[2]], m[s[3]])
        B2G(2, 6, 10, 14, m[s[4]], m[s[5]]) B2G(3, 7, 11, 15, m[s[6]], m[s[7]])
        B2G(0, 5, 10, 15, m[s[8]], m[s[9]]) B2G(1, 6, 11, 12, m[s[10]], m[s[11]])
        B2G(2, 7, 8, 13, m[s[12]], m[s[13]]) B2G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef B2G
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
    memcpy(out, h, (size_t)outlen);
}

// ---------------------------------------------------------------- HMAC-SHA256 with cached pad states
struct HmacKey {
    SHA256_CTX inner, outer;
    void init(const uint8_t* key, size_t klen)
    {
        uint8_t k[64] = {0}, pad[64];
        if (klen > 64) SHA256(key, klen, k); else memcpy(k, key, klen);
        for (int i = 0; i < 64; ++i) pad[i] = k[i] ^ 0x36;
        SHA256_Init(&inner); SHA256_Update(&inner, pad, 64);
        for (int i = 0; i < 64; ++i) pad[i] = k[i] ^ 0x5c;
        SHA256_Init(&outer); SHA256_Update(&outer, pad, 64);
    }
    void mac(const uint8_t* msg, size_t len, uint8_t out[32]) const
    {
        SHA256_CTX c = inner;
        uint8_t d[32];
        SHA256_Update(&c, msg, len); SHA256_Final(d, &c);
        c = outer;
        SHA256_Update(&c, d, 32); SHA256_Final(out, &c);
    }
};

struct KeyCtx {
    uint8_t aead_key[32];
    uint8_t prng_sub[16];
    HmacKey band;                 // band key = the raw 32-byte master key (rtwm/detector.py:31)
    uint8_t hdr_pn[16];           // pn_bits(0, 128) packed
    std::vector<uint8_t> hop;     // band index per counter, grown on demand
};

struct Feeder {
    std::vector<KeyCtx> keys;
    int nthreads;
};

// Process-wide worker pool: the feeder is called a few times per sub-batch (and once per block by the live TX
// service), so spawning threads per call costs more than the work of a small call.  Workers are created on demand,
// detached and never torn down (no static-destruction order to get wrong); one parallel region runs at a time.
class WorkPool {
public:
    // one pool per PROCESS: after fork() the child owns no worker threads (only the forking thread survives), so a pool
    // inherited from the parent would wait for helpers that do not exist; the child gets a fresh one (the old is leaked)
    static WorkPool& get()
    {
        static std::atomic<WorkPool*> p{nullptr};
        static std::atomic<long> owner{0};
        const long me = (long)getpid();
        if (owner.load() != me || !p.load()) { p.store(new WorkPool()); owner.store(me); }
        return *p.load();
    }
    void run(int n, int nthreads, const std::function<void(int)>& fn)
    {
        std::lock_guard<std::mutex> region(region_mu_);
        const int helpers = nthreads - 1;
        while ((int)nworkers_ < helpers) {
            const int id = nworkers_++;
            std::thread([this, id]() { worker(id); }).detach();
        }
        fn_ = &fn; n_ = n; next_.store(0);
        {
            std::lock_guard<std::mutex> lk(mu_);
            want_ = helpers; active_ = helpers; ++gen_;
        }
        cv_start_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return active_ == 0; });
    }
private:
    void work() { for (;;) { const int i = next_.fetch_add(1); if (i >= n_) break; (*fn_)(i); } }
    void worker(int id)
    {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_start_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            const bool part = id < want_;
            lk.unlock();
            if (!part) continue;
            work();
            lk.lock();
            if (--active_ == 0) cv_done_.notify_one();
        }
    }
    std::mutex region_mu_, mu_;
    std::condition_variable cv_start_, cv_done_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_ = 0, want_ = 0, active_ = 0, nworkers_ = 0;
    unsigned long long gen_ = 0;
    std::atomic<int> next_{0};
};

template <class F> static void parallel_for(int n, int nthreads, F fn)
{
    if (n <= 0) return;
    nthreads = std::max(1, std::min(nthreads, n));
    if (nthreads == 1) { for (int i = 0; i < n; ++i) fn(i); return; }
    const std::function<void(int)> f = [&](int i) { fn(i); };
    WorkPool::get().run(n, nthreads, f);
}

// OpenSSL 3: EVP_aes_128_ecb() / EVP_chacha20_poly1305() make every *Init_ex do an implicit algorithm fetch under a
// library-wide lock, which serialises the feeder's threads.  Fetch both once and hand the fetched objects out.
static const EVP_CIPHER* cipher_aes_ecb()
{
    static EVP_CIPHER* c = EVP_CIPHER_fetch(nullptr, "AES-128-ECB", nullptr);
    return c ? c : EVP_aes_128_ecb();
}
static const EVP_CIPHER* cipher_chacha()
{
    static EVP_CIPHER* c = EVP_CIPHER_fetch(nullptr, "ChaCha20-Poly1305", nullptr);
    return c ? c : EVP_chacha20_poly1305();
}

// AES-128-ECB over counter blocks (ctr << 64 | blk), big-endian (rtwm/utils.py:115-124)
struct AesEcb {
    EVP_CIPHER_CTX* ctx;
    explicit AesEcb(const uint8_t key[16])
    {
        ctx = EVP_CIPHER_CTX_new();
        EVP_EncryptInit_ex(ctx, cipher_aes_ecb(), nullptr, key, nullptr);
        EVP_CIPHER_CTX_set_padding(ctx, 0);
    }
    // a context bound to AES-128-ECB once; rekey() only expands a key (no reference traffic on the shared cipher object)
    AesEcb()
    {
        ctx = EVP_CIPHER_CTX_new();
        EVP_EncryptInit_ex(ctx, cipher_aes_ecb(), nullptr, nullptr, nullptr);
        EVP_CIPHER_CTX_set_padding(ctx, 0);
    }
    void rekey(const uint8_t key[16]) { EVP_EncryptInit_ex(ctx, nullptr, nullptr, key, nullptr); }
    ~AesEcb() { EVP_CIPHER_CTX_free(ctx); }
    void pn(uint64_t ctr, uint8_t out[PN_BYTES])
    {
        uint8_t in[16 * PN_BLOCKS], enc[16 * PN_BLOCKS];
        for (int b = 0; b < PN_BLOCKS; ++b) {
            for (int i = 0; i < 8; ++i) in[16 * b + i] = (uint8_t)(ctr >> (56 - 8 * i));
            for (int i = 0; i < 8; ++i) in[16 * b + 8 + i] = (uint8_t)((uint64_t)b >> (56 - 8 * i));
        }
        int ol = 0;
        EVP_EncryptUpdate(ctx, enc, &ol, in, sizeof(in));
        memcpy(out, enc, PN_BYTES);
    }
};

static void derive_key(const uint8_t key32[32], KeyCtx& k)
{
    // HKDF-SHA256, salt = None (32 zero bytes), info = "EchoSeal:KDF:v1", 64 bytes (rtwm/crypto.py:19-27)
    static const uint8_t info[] = "EchoSeal:KDF:v1";
    const size_t ilen = sizeof(info) - 1;
    uint8_t salt[32] = {0}, prk[32], t1[32], t2[32], buf[32 + 32];
    HmacKey ext, exp;                       // (the one-shot HMAC() fetches SHA-256 under a library lock per call)
    ext.init(salt, 32);
    ext.mac(key32, 32, prk);
    exp.init(prk, 32);
    memcpy(buf, info, ilen); buf[ilen] = 1;
    exp.mac(buf, ilen + 1, t1);
    memcpy(buf, t1, 32); memcpy(buf + 32, info, ilen); buf[32 + ilen] = 2;
    exp.mac(buf, 32 + ilen + 1, t2);
    memcpy(k.aead_key, t1, 32);
    // StreamPRNG sub-key: BLAKE2s-128(prng_key, person="EchoSeal") (rtwm/utils.py:94)
    static const uint8_t person[8] = {'E', 'c', 'h', 'o', 'S', 'e', 'a', 'l'};
    blake2s_short(t2, 32, person, k.prng_sub, 16);
    k.band.init(key32, 32);
    AesEcb aes(k.prng_sub);
    uint8_t pn[PN_BYTES];
    aes.pn(0, pn);
    memcpy(k.hdr_pn, pn, 16);
    k.hop.clear();
}

static void grow_hop(KeyCtx& k, size_t hi)
{
    if (k.hop.size() >= hi) return;
    const size_t lo = k.hop.size();
    const size_t nhi = std::max(hi, std::max<size_t>(512, 2 * lo));
    k.hop.resize(nhi);
    uint8_t msg[4], d[32];
    for (size_t c = lo; c < nhi; ++c) {
        msg[0] = (uint8_t)(c >> 24); msg[1] = (uint8_t)(c >> 16); msg[2] = (uint8_t)(c >> 8); msg[3] = (uint8_t)c;
        k.band.mac(msg, 4, d);
        k.hop[c] = d[0] & 3;                 // digest()[0] % 4 (rtwm/utils.py:33-36)
    }
}

// A ChaCha20-Poly1305 context bound to its cipher ONCE (encrypt or decrypt side); every message then only sets key and
// nonce.  Passing the cipher to EVP_*Init_ex per message takes and drops a reference on the shared cipher object - an
// atomic on one cache line from every worker thread, which is what kept the feeder from scaling past a few threads.
struct Aead {
    EVP_CIPHER_CTX* ctx = nullptr;
    int enc = -1;
    ~Aead() { if (ctx) EVP_CIPHER_CTX_free(ctx); }
    bool bind(int want_enc)
    {
        if (ctx && enc == want_enc) return true;
        if (!ctx) ctx = EVP_CIPHER_CTX_new();
        if (EVP_CipherInit_ex(ctx, cipher_chacha(), nullptr, nullptr, nullptr, want_enc) != 1) return false;
        EVP_CIPHER_CTX_ctrl(ctx, EVP_CTRL_AEAD_SET_IVLEN, 12, nullptr);
        enc = want_enc;
        return true;
    }
};

// ChaCha20-Poly1305 open of nonce(12) | ct(27) | tag(16) (rtwm/crypto.py:39-43)
static bool aead_open(Aead& A, const uint8_t key[32], const uint8_t blob[55], uint8_t pt[27])
{
    int ol = 0, fl = 0;
    if (!A.bind(0)) return false;
    EVP_CIPHER_CTX* ctx = A.ctx;
    if (EVP_DecryptInit_ex(ctx, nullptr, nullptr, key, blob) != 1) return false;
    if (EVP_DecryptUpdate(ctx, pt, &ol, blob + 12, 27) != 1) return false;
    EVP_CIPHER_CTX_ctrl(ctx, EVP_CTRL_AEAD_SET_TAG, 16, (void*)(blob + 39));
    return EVP_DecryptFinal_ex(ctx, pt + ol, &fl) == 1;
}

static bool aead_seal(Aead& A, const uint8_t key[32], const uint8_t nonce[12], const uint8_t pt[27], uint8_t blob[55])
{
    int ol = 0, fl = 0;
    if (!A.bind(1)) return false;
    EVP_CIPHER_CTX* ctx = A.ctx;
    if (EVP_EncryptInit_ex(ctx, nullptr, nullptr, key, nonce) != 1) return false;
    memcpy(blob, nonce, 12);
    if (EVP_EncryptUpdate(ctx, blob + 12, &ol, pt, 27) != 1) return false;
    if (EVP_EncryptFinal_ex(ctx, blob + 12 + ol, &fl) != 1) return false;
    EVP_CIPHER_CTX_ctrl(ctx, EVP_CTRL_AEAD_GET_TAG, 16, blob + 39);
    return true;
}

}  // namespace

extern "C" {

// ---- key bank -------------------------------------------------------------------------------
void* es_host_keys_new(const uint8_t* keys /*[nkeys][32]*/, int nkeys, int nthreads)
{
    Feeder* f = new Feeder();
    f->keys.resize((size_t)std::max(0, nkeys));
    if (nthreads <= 0) {
        // default: the host cores divided among the ranks of this node (torchrun exports LOCAL_WORLD_SIZE), so that N
        // ranks on one box do not each start a full set of workers
        int ranks = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
        nthreads = (int)std::max(1u, std::thread::hardware_concurrency() / (unsigned)ranks);
    }
    f->nthreads = nthreads;
    parallel_for(nkeys, f->nthreads, [&](int i) { derive_key(keys + 32 * (size_t)i, f->keys[(size_t)i]); });
    return f;
}

void es_host_keys_free(void* h)
{
    Feeder* f = (Feeder*)h;
    if (!f) return;
    delete f;
}

int es_host_threads(void* h) { return ((Feeder*)h)->nthreads; }

int es_host_hdr_pn(void* h, const int32_t* key_idx, int n, uint8_t* out /*[n][16]*/)
{
    Feeder* f = (Feeder*)h;
    for (int i = 0; i < n; ++i) {
        const int k = key_idx ? key_idx[i] : i;
        if (k < 0 || (size_t)k >= f->keys.size()) return -1;
        memcpy(out + 16 * (size_t)i, f->keys[(size_t)k].hdr_pn, 16);
    }
    return 0;
}

int es_host_hop(void* h, int key, uint32_t lo, uint32_t hi, uint8_t* out)
{
    Feeder* f = (Feeder*)h;
    if (key < 0 || (size_t)key >= f->keys.size() || hi < lo) return -1;
    grow_hop(f->keys[(size_t)key], hi);
    memcpy(out, f->keys[(size_t)key].hop.data() + lo, hi - lo);
    return 0;
}

int es_host_pn(void* h, int key, const uint64_t* ctrs, int n, uint8_t* out /*[n][152]*/)
{
    Feeder* f = (Feeder*)h;
    if (key < 0 || (size_t)key >= f->keys.size()) return -1;
    AesEcb aes(f->keys[(size_t)key].prng_sub);
    for (int i = 0; i < n; ++i) aes.pn(ctrs[i], out + (size_t)PN_BYTES * i);
    return 0;
}

// ---- RX: candidate counters + budget + PN bits (rtwm/detector.py:105-151) ----------------------
// Pass 1 (counts only: item_* == NULL) and pass 2 (fill) are both driven by this one routine so the two
// can never disagree.  Items are stored clip-major, band-major (BAND_PLAN order), attempt order.
//   peaks i32[nb][4][25], npeaks i32[nb][4], hdr f32[nb][4][25][4] (ok(-1 = no frame), val, score, margin)
//   band_count i32[nb][4] (out), item_offset i64[nb+1] (in for pass 2 / out for pass 1)
int64_t es_host_rx_enumerate(void* h, const int32_t* key_idx, int nb, int n_samples,
                             const int32_t* peaks, const int32_t* npeaks, const float* hdr,
                             int32_t* band_count, int64_t* item_offset,
                             int32_t* item_peak, uint32_t* item_ctr, int32_t* item_clip, uint8_t* pn)
{
    Feeder* f = (Feeder*)h;
    const bool fill = item_peak != nullptr;
    std::atomic<int> bad(0);
    {   // hop tables are per key and not thread-safe to grow: size them up front, one task per distinct key
        const size_t hi = (size_t)(n_samples / FRAME_LEN) + WIDE + 4;
        std::vector<int> ks;
        ks.reserve((size_t)nb);
        for (int ci = 0; ci < nb; ++ci) {
            const int kidx = key_idx ? key_idx[ci] : ci;
            if (kidx < 0 || (size_t)kidx >= f->keys.size()) return -1;
            if (f->keys[(size_t)kidx].hop.size() < hi) ks.push_back(kidx);
        }
        std::sort(ks.begin(), ks.end());
        ks.erase(std::unique(ks.begin(), ks.end()), ks.end());
        parallel_for((int)ks.size(), f->nthreads, [&](int q) { grow_hop(f->keys[(size_t)ks[(size_t)q]], hi); });
    }
    parallel_for(nb, f->nthreads, [&](int ci) {
        const int kidx = key_idx ? key_idx[ci] : ci;
        if (kidx < 0 || (size_t)kidx >= f->keys.size()) { bad = 1; return; }
        KeyCtx& k = f->keys[(size_t)kidx];
        int64_t pos = fill ? item_offset[ci] : 0;
        int64_t total = 0;
        AesEcb* aes = fill ? new AesEcb(k.prng_sub) : nullptr;
        for (int bi = 0; bi < 4; ++bi) {
            int tried = 0;
            bool stop = false;
            const int np = npeaks[ci * 4 + bi];
            for (int slot = 0; slot < np && slot < PEAK_LIMIT && !stop; ++slot) {
                const size_t pidx = ((size_t)ci * 4 + (size_t)bi) * PEAK_LIMIT + (size_t)slot;
                const int start = peaks[pidx];
                if (start < 0 || start + FRAME_LEN > n_samples) continue;
                const bool ok = hdr[pidx * 4 + 0] > 0.5f;
                const int val = (int)hdr[pidx * 4 + 1];
                const long est = lrint((double)start / (double)FRAME_LEN);      // Python round(): half to even
                const long lo = std::max(0L, est - WIDE), hi = est + WIDE + 1;
                if ((size_t)hi > k.hop.size()) { bad = 1; return; }
                auto emit = [&](long c) {
                    if (fill) {
                        item_peak[pos] = (int32_t)pidx;
                        item_ctr[pos] = (uint32_t)c;
                        item_clip[pos] = ci;
                        aes->pn((uint64_t)c, pn + (size_t)PN_BYTES * (size_t)pos);
                        ++pos;
                    }
                    ++total; ++tried;
                    if (tried >= MAX_TRIES) stop = true;
                };
                if (ok) {
                    for (long c = lo; c < hi && !stop; ++c)
                        if ((int)(c & 0xFFFF) == val && k.hop[(size_t)c] == bi) emit(c);
                } else {
                    const long tl = std::max(0L, est - TIGHT), th = est + TIGHT + 1;
                    int any = 0;
                    for (long c = tl; c < th; ++c) any += (k.hop[(size_t)c] == bi);
                    if (any) { for (long c = tl; c < th && !stop; ++c) if (k.hop[(size_t)c] == bi) emit(c); }
                    else { for (long c = lo; c < hi && !stop; ++c) if (k.hop[(size_t)c] == bi) emit(c); }
                }
            }
            band_count[ci * 4 + bi] = tried;
        }
        delete aes;
        if (!fill) item_offset[ci + 1] = total;      // per-clip count; prefix-summed below
    });
    if (bad) return -1;
    if (!fill) {
        item_offset[0] = 0;
        for (int ci = 0; ci < nb; ++ci) item_offset[ci + 1] += item_offset[ci];
    }
    return item_offset[nb];
}

// ---- RX: AEAD validation of the CRC-passing candidates, in the reference's order ----------------
// hits are (codeword, slot, payload) with codeword = 4*item + variant, slot 0 = hard decision, 1.. = list
// rank + 1; they must be sorted by (codeword, slot).  Per clip: hop-0 band first, then the other bands in
// BAND_PLAN order; within a band the attempts in order; per attempt variants 0..3; per variant slots
// ascending (rtwm/detector.py:44-53, 144-151, 161-233; rtwm/fastpolar.py:269-276, 335-349).
//   nonce_state u8[nb][9]: [0] = has session nonce, [1..8] = nonce (in/out; the latch of :223-233)
int es_host_rx_validate(void* h, const int32_t* key_idx, int nb,
                        const int32_t* band_count, const int64_t* item_offset, const uint32_t* item_ctr,
                        const int64_t* hit_cw, const int32_t* hit_slot, const uint8_t* hit_payload, int64_t nhits,
                        uint8_t* nonce_state, uint8_t* verdict /*[nb]*/, uint8_t* plaintext /*[nb][27]*/)
{
    Feeder* f = (Feeder*)h;
    (void)hit_slot;
    for (int ci = 0; ci < nb; ++ci) {               // hop(0) is needed for the band order; normally already there
        const int kidx = key_idx ? key_idx[ci] : ci;
        if (kidx < 0 || (size_t)kidx >= f->keys.size()) return -1;
        if (f->keys[(size_t)kidx].hop.empty()) grow_hop(f->keys[(size_t)kidx], 1);
    }
    parallel_for(nb, f->nthreads, [&](int ci) {
        const int kidx = key_idx ? key_idx[ci] : ci;
        KeyCtx& k = f->keys[(size_t)kidx];
        verdict[ci] = 0;
        const int hop0 = k.hop[0];
        int order[4] = {hop0, 0, 0, 0};
        for (int b = 0, o = 1; b < 4; ++b) if (b != hop0) order[o++] = b;
        int64_t boff[4];
        int64_t o = item_offset[ci];
        for (int b = 0; b < 4; ++b) { boff[b] = o; o += band_count[ci * 4 + b]; }
        Aead ctx;
        uint8_t* ns = nonce_state + 9 * (size_t)ci;
        for (int oi = 0; oi < 4 && !verdict[ci]; ++oi) {
            const int b = order[oi];
            for (int a = 0; a < band_count[ci * 4 + b] && !verdict[ci]; ++a) {
                const int64_t it = boff[b] + a;
                const int64_t cw0 = 4 * it, cw1 = 4 * it + 4;
                const int64_t* p0 = std::lower_bound(hit_cw, hit_cw + nhits, cw0);
                const int64_t* p1 = std::lower_bound(hit_cw, hit_cw + nhits, cw1);
                for (const int64_t* p = p0; p < p1; ++p) {
                    const uint8_t* blob = hit_payload + 55 * (size_t)(p - hit_cw);
                    uint8_t pt[27 + 16];
                    if (!aead_open(ctx, k.aead_key, blob, pt)) continue;
                    if (memcmp(pt, "ESAL", 4) != 0) continue;
                    const uint32_t ec = ((uint32_t)pt[4] << 24) | ((uint32_t)pt[5] << 16) | ((uint32_t)pt[6] << 8) | pt[7];
                    if (ec != item_ctr[it]) continue;
                    // first valid candidate of this attempt decides it (polar_dec returns it); nonce latch
                    if (!ns[0] || memcmp(ns + 1, pt + 8, 8) == 0) {
                        ns[0] = 1; memcpy(ns + 1, pt + 8, 8);
                        verdict[ci] = 1;
                        if (plaintext) memcpy(plaintext + 27 * (size_t)ci, pt, 27);
                    }
                    break;
                }
            }
        }
    });
    return 0;
}

// ---- TX: sealed payloads, PN bits, hop bands for a batch of frames (rtwm/embedder.py:82-119,153-168) ----
//   rnd u8[F][23] = 11 bytes of padding + 12-byte AEAD nonce per frame (the caller's randomness)
int es_host_tx_prepare(void* h, const int32_t* key_idx, const uint32_t* ctr, const uint8_t* session_nonce /*[F][8]*/,
                       const uint8_t* rnd, int F, uint8_t* payload /*[F][55]*/, uint8_t* pn /*[F][152]*/,
                       uint8_t* hdr_pn /*[F][16]*/, int32_t* band, int32_t* ctr_lo16)
{
    Feeder* f = (Feeder*)h;
    std::atomic<int> bad(0);
    const int chunk = 64;
    const int nchunks = (F + chunk - 1) / chunk;
    parallel_for(nchunks, f->nthreads, [&](int cidx) {
        Aead ctx;                                     // one AEAD and one AES context per chunk of frames
        AesEcb aes;
        int last_k = -1;
        const int i1 = std::min(F, (cidx + 1) * chunk);
        for (int i = cidx * chunk; i < i1; ++i) {
            const int kidx = key_idx ? key_idx[i] : 0;
            if (kidx < 0 || (size_t)kidx >= f->keys.size()) { bad = 1; continue; }
            const KeyCtx& k = f->keys[(size_t)kidx];
            if (kidx != last_k) { aes.rekey(k.prng_sub); last_k = kidx; }
            uint8_t meta[27];
            memcpy(meta, "ESAL", 4);
            meta[4] = (uint8_t)(ctr[i] >> 24); meta[5] = (uint8_t)(ctr[i] >> 16); meta[6] = (uint8_t)(ctr[i] >> 8); meta[7] = (uint8_t)ctr[i];
            memcpy(meta + 8, session_nonce + 8 * (size_t)i, 8);
            memcpy(meta + 16, rnd + 23 * (size_t)i, 11);
            if (!aead_seal(ctx, k.aead_key, rnd + 23 * (size_t)i + 11, meta, payload + 55 * (size_t)i)) bad = 1;
            aes.pn((uint64_t)ctr[i], pn + (size_t)PN_BYTES * (size_t)i);
            memcpy(hdr_pn + 16 * (size_t)i, k.hdr_pn, 16);
            uint8_t msg[4] = {meta[4], meta[5], meta[6], meta[7]}, d[32];
            k.band.mac(msg, 4, d);
            band[i] = d[0] & 3;
            ctr_lo16[i] = (int32_t)(ctr[i] & 0xFFFF);
        }
    });
    return bad ? -1 : 0;
}

}  // extern "C"
