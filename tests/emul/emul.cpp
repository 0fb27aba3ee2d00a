/*
 * emul.cpp — TEST INFRASTRUCTURE: host-compiled harness around the device
 * composers of registrar_b200/csrc/regk_core.cuh.
 *
 * There is no GPU in the build container, so the word-wise (SWAR) logic that the
 * kernels run per thread is exercised here under g++ on the same inputs the
 * oracle sees: the staging layout (16-byte phase of the tile in the packed
 * stream), the shared-word WordSink protocol between neighbouring records, the
 * length formulas and the fence.  It mirrors the per-tile control flow of
 * regk_kernels.cuh with a serial scan in place of the look-back.  Built and
 * used only by tests/test_core_emul.py; it is not part of the product library.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/regk.h"
#include "../../registrar_b200/csrc/regk_core.cuh"
#include "../../registrar_b200/csrc/regk_decode_core.cuh"

using namespace regk;

namespace {
constexpr uint32_t TILE = 128;          /* as REGK_TILE in regk_kernels.cuh */

struct Frags {
    std::vector<uint8_t> blob;
};
}

extern "C" {

/* returns OR of bad bits; outputs must be large enough (caller sizes from the oracle) */
uint32_t emul_paths(const regk_batch *b, int generic, uint8_t *out_bytes, uint64_t *out_off, uint64_t *first_bad)
{
    const bool alias = b->flags & REGK_NODE_ALIAS;
    const uint64_t n = b->n;
    uint32_t bad_all = 0;
    uint64_t fb = UINT64_MAX;
    uint64_t base = 0;
    for (uint64_t r0 = 0; r0 < n; r0 += TILE) {
        const uint32_t nrec = (uint32_t)std::min<uint64_t>(TILE, n - r0);
        const uint64_t D0 = b->domain_off[r0], D1 = b->domain_off[r0 + nrec];
        const uint64_t HB0 = alias ? 0 : (b->host_off ? b->host_off[r0] : r0 * b->host_stride);
        const uint64_t HB1 = alias ? 0 : (b->host_off ? b->host_off[r0 + nrec] : (r0 + nrec) * b->host_stride);
        const uint64_t da0 = D0 & ~15ull, ha0 = HB0 & ~15ull;
        /* staged images, poisoned outside the tile's own bytes */
        /* 16 bytes of front padding in front of the staged domain bytes, like the kernel's s_dom */
        std::vector<uint32_t> sdom_buf((D1 - da0) / 4 + 20, 0xA5A5A5A5u), shost((HB1 - ha0) / 4 + 16, 0x5A5A5A5Au);
        uint32_t *sdom_p = sdom_buf.data() + 4;
        struct { uint32_t *p; uint32_t *data() { return p; } } sdom{sdom_p};
        if (D1 > da0)
            memcpy(sdom.data(), b->domain_bytes + da0, D1 - da0);
        if (HB1 > ha0)
            memcpy(shost.data(), b->host_bytes + ha0, HB1 - ha0);
        std::vector<uint32_t> len(nrec), local(nrec);
        std::vector<DomainStats> st(nrec);
        std::vector<DomainInfo> di(nrec);
        std::vector<uint32_t> sbits((D1 - da0) / 32 + 16, 0xC3C3C3C3u);
        uint32_t suspicious = 0;
        if (!generic) {
            /* cooperative pre-pass, "threads" in a scrambled order */
            for (uint32_t tt = 0; tt < TILE; tt++) {
                const uint32_t t = (tt * 37u) % TILE;
                suspicious |= prepass_domain(sdom.data(), (uint16_t *)sbits.data(), (uint32_t)((D1 - da0 + 15) >> 4), t, TILE);
                if (!alias)
                    suspicious |= prepass_host(shost.data(), (uint32_t)((HB1 - ha0 + 15) >> 4), t, TILE);
            }
        }
        uint32_t total = 0;
        for (uint32_t t = 0; t < nrec; t++) {
            const uint64_t r = r0 + t;
            const uint32_t d0 = b->domain_off[r], L = b->domain_off[r + 1] - d0;
            const uint64_t h0 = alias ? 0 : (b->host_off ? b->host_off[r] : r * b->host_stride);
            const uint32_t H = alias ? 0 : (b->host_off ? b->host_off[r + 1] - b->host_off[r] : b->host_stride);
            uint32_t bad = 0;
            if (generic) {
                GuardedWords ds{(const uint32_t *)b->domain_bytes}, hs{(const uint32_t *)b->host_bytes};
                st[t] = scan_domain(ds, d0, L);
                bad = st[t].bad | (alias ? 0 : check_host(hs, (uint32_t)h0, H));
                len[t] = path_length(st[t], L, H, alias);
            } else {
                PaddedWords ds{sdom.data()}, hs{shost.data()};
                di[t] = domain_info(sbits.data(), (uint32_t)(d0 - da0), L);
                if (suspicious) {
                    bad = recheck_domain((const uint8_t *)sdom.data(), sbits.data(), (uint32_t)(d0 - da0), L);
                    if (!alias)
                        bad |= check_host(hs, (uint32_t)(h0 - ha0), H);
                } else if (!alias && H <= 2) {
                    bad = check_host(hs, (uint32_t)(h0 - ha0), H);
                }
                len[t] = path_length2(di[t], L, H, alias);
            }
            if (bad) {
                bad_all |= bad;
                fb = std::min<uint64_t>(fb, r);
            }
            local[t] = total;
            total += len[t];
        }
        const uint32_t shift = (uint32_t)(base & 15);
        std::vector<uint32_t> sout((shift + total) / 4 + 16, 0xEEEEEEEEu);
        std::vector<WordSink> sinks(nrec);
        /* compose in reverse thread order: neighbours must not clobber shared words */
        for (uint32_t tt = nrec; tt-- > 0;) {
            const uint64_t r = r0 + tt;
            const uint32_t d0 = b->domain_off[r], L = b->domain_off[r + 1] - d0;
            const uint64_t h0 = alias ? 0 : (b->host_off ? b->host_off[r] : r * b->host_stride);
            const uint32_t H = alias ? 0 : (b->host_off ? b->host_off[r + 1] - b->host_off[r] : b->host_stride);
            out_off[r] = base + local[tt];
            if (generic) {
                GuardedWords ds{(const uint32_t *)b->domain_bytes}, hs{(const uint32_t *)b->host_bytes};
                ByteSink sink;
                sink.init(out_bytes + base + local[tt]);
                if (alias)
                    emit_path<true>(ds, d0, L, hs, (uint32_t)h0, H, sink);
                else
                    emit_path<false>(ds, d0, L, hs, (uint32_t)h0, H, sink);
                if ((uint64_t)(sink.p - (out_bytes + base + local[tt])) != len[tt])
                    abort();
            } else {
                PaddedWords ds{sdom.data()}, hs{shost.data()};
                WordSink &sink = sinks[tt];
                sink.init(sout.data(), local[tt] + shift);
                const uint32_t dof = (uint32_t)(d0 - da0), hof = (uint32_t)(h0 - ha0);
                if (alias)
                    emit_path2<true, false>(sdom.data(), sbits.data(), dof, L, di[tt], shost.data(), hof, H, sink);
                else if (H >= 24)
                    emit_path2<false, true>(sdom.data(), sbits.data(), dof, L, di[tt], shost.data(), hof, H, sink);
                else
                    emit_path2<false, false>(sdom.data(), sbits.data(), dof, L, di[tt], shost.data(), hof, H, sink);
            }
        }
        if (!generic)
            for (uint32_t tt = 0; tt < nrec; tt++)     /* phase B after the "barrier" */
                sinks[tt].tail();
        if (!generic)
            memcpy(out_bytes + base, (const uint8_t *)sout.data() + shift, total);
        base += total;
    }
    out_off[n] = base;
    if (first_bad)
        *first_bad = fb;
    return bad_all;
}

/* frag blob: same layout regk_set_types builds (TypeFrag[] then word-aligned fragments) */
uint32_t emul_jsons(const regk_batch *b, const uint8_t *blob, uint32_t ntypes, int generic, uint8_t *out_bytes,
    uint64_t *out_off, uint64_t *first_bad)
{
    const uint64_t n = b->n;
    uint32_t bad_all = 0;
    uint64_t fb = UINT64_MAX;
    uint64_t base = 0;
    const PaddedWords fsrc{(const uint32_t *)blob};
    const GuardedWords asrc{(const uint32_t *)b->addr_bytes};
    for (uint64_t r0 = 0; r0 < n; r0 += TILE) {
        const uint32_t nrec = (uint32_t)std::min<uint64_t>(TILE, n - r0);
        struct Rec {
            TypeFrag tf;
            uint32_t aw[4];
            uint32_t a0, al, k, p0;
            int32_t ttl;
            bool has_ttl, has_ports;
            uint32_t len, local;
        };
        std::vector<Rec> rec(nrec);
        uint32_t total = 0;
        for (uint32_t t = 0; t < nrec; t++) {
            const uint64_t r = r0 + t;
            Rec &q = rec[t];
            uint32_t bad = 0;
            uint32_t tid = b->type_id[r];
            if (tid >= ntypes) {
                bad |= BAD_TYPE_ID;
                tid = 0;
            }
            q.tf = ((const TypeFrag *)blob)[tid];
            q.a0 = b->addr_off[r];
            q.al = b->addr_off[r + 1] - q.a0;
            if (q.al == 0)
                bad |= BAD_ADDR_BYTE;
            q.ttl = b->ttl ? b->ttl[r] : INT32_MIN;
            q.has_ttl = q.ttl != INT32_MIN;
            q.p0 = 0;
            q.k = 0;
            if (b->ports_off) {
                q.p0 = b->ports_off[r];
                q.k = b->ports_off[r + 1] - q.p0;
            }
            q.has_ports = b->ports_present ? b->ports_present[r] != 0 : q.k > 0;
            if (!q.has_ports)
                q.k = 0;
            memset(q.aw, 0, sizeof q.aw);
            {
                const uint32_t n16 = q.al < 16u ? q.al : 16u;
                const uint32_t sh = (q.a0 & 3u) * 8u;
                uint32_t wi = q.a0 >> 2;
                uint32_t lo = n16 ? asrc.word(wi) : 0u;
                for (uint32_t w = 0; w < 4; w++) {
                    if (n16 > 4u * w) {
                        const uint32_t nb = std::min(4u, n16 - 4u * w);
                        const uint32_t hi = asrc.word_hi(wi + 1, sh + 8u * nb > 32u || n16 > 4u * (w + 1));
                        const uint32_t keep = low_bytes(nb);
                        q.aw[w] = funnel_r(lo, hi, sh) & keep;
                        if (addr_word_bad(q.aw[w], keep))
                            bad |= BAD_ADDR_BYTE;
                        lo = hi;
                        wi++;
                    }
                }
                for (uint32_t i = q.a0 + 16u; i < q.a0 + q.al; i++) {
                    const uint32_t c = b->addr_bytes[i];
                    if (c < 0x20u || c >= 0x80u || c == 0x22u || c == 0x5Cu)
                        bad |= BAD_ADDR_BYTE;
                }
            }
            uint32_t pd = 0;
            for (uint32_t i = 0; i < q.k; i++)
                pd += ndigits_u32(b->ports[q.p0 + i]);
            q.len = json_length(q.tf.f1_len, q.tf.f2_len, q.al, q.has_ttl, q.ttl, q.has_ports, q.k, pd);
            q.local = total;
            total += q.len;
            if (bad) {
                bad_all |= bad;
                fb = std::min<uint64_t>(fb, r);
            }
        }
        const uint32_t shift = (uint32_t)(base & 15);
        std::vector<uint32_t> sout((shift + total) / 4 + 16, 0xEEEEEEEEu);
        std::vector<WordSink> sinks(nrec);
        for (uint32_t tt = nrec; tt-- > 0;) {
            const uint64_t r = r0 + tt;
            Rec &q = rec[tt];
            out_off[r] = base + q.local;
            const uint32_t *ports = b->ports + q.p0;
            auto port = [ports](uint32_t i) { return ports[i]; };
            if (generic) {
                ByteSink sink;
                sink.init(out_bytes + base + q.local);
                emit_json(fsrc, q.tf, q.aw, asrc, q.a0, q.al, q.has_ttl, q.ttl, q.has_ports, q.k, port, sink);
                if ((uint64_t)(sink.p - (out_bytes + base + q.local)) != q.len)
                    abort();
            } else {
                WordSink &sink = sinks[tt];
                sink.init(sout.data(), q.local + shift);
                emit_json(fsrc, q.tf, q.aw, asrc, q.a0, q.al, q.has_ttl, q.ttl, q.has_ports, q.k, port, sink);
            }
        }
        if (!generic) {
            for (uint32_t tt = 0; tt < nrec; tt++)
                sinks[tt].tail();
            memcpy(out_bytes + base, (const uint8_t *)sout.data() + shift, total);
        }
        base += total;
    }
    out_off[n] = base;
    if (first_bad)
        *first_bad = fb;
    return bad_all;
}

uint32_t emul_dec(uint32_t v, uint8_t *out)
{
    ByteSink s;
    s.init(out);
    put_u32_dec(v, s);
    return (uint32_t)(s.p - out);
}

uint32_t emul_ndigits(uint32_t v)
{
    return ndigits_u32(v);
}

/* setupDirectories pass (regk_parents.cuh) run serially with the same helpers: directory lengths, table
   insertion with byte-exact comparison, first occurrences in record order.  `path_words` must be readable one
   word past the stream.  tail_mode as in ParentParams.  Returns the number of distinct directories. */
uint64_t emul_parents(const uint32_t *path_words, const uint64_t *path_off, uint64_t n, uint32_t tail_mode,
    uint32_t host_stride, const uint32_t *host_off, uint32_t *parent_len, uint64_t *unique_first)
{
    const uint8_t *bytes = reinterpret_cast<const uint8_t *>(path_words);
    uint64_t slots = 16;                                        /* small on purpose: probing gets exercised */
    while (slots < 2 * n)
        slots <<= 1;
    std::vector<uint32_t> owner(slots, 0), first(slots, 0xFFFFFFFFu), slot_of(n);
    auto plen_of = [&](uint64_t i) -> uint32_t {
        const uint32_t len = (uint32_t)(path_off[i + 1] - path_off[i]);
        if (len == 0)
            return 0;
        if (tail_mode == 1)
            return dirname_len_host(len, host_stride);
        if (tail_mode == 2)
            return dirname_len_host(len, host_off[i + 1] - host_off[i]);
        return dirname_len_scan(bytes + path_off[i], len);
    };
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t plen = plen_of(i);
        parent_len[i] = plen;
        uint32_t slot = string_hash32(path_words, path_off[i], plen) & (uint32_t)(slots - 1);
        for (;;) {
            if (owner[slot] == 0) {
                owner[slot] = (uint32_t)i + 1;
                break;
            }
            const uint64_t j = owner[slot] - 1;
            if (plen_of(j) == plen && string_equal(path_words, path_off[i], path_off[j], plen))
                break;
            slot = (slot + 1) & (uint32_t)(slots - 1);
        }
        slot_of[i] = slot;
        if (first[slot] > (uint32_t)i)
            first[slot] = (uint32_t)i;
    }
    uint64_t nu = 0;
    for (uint64_t i = 0; i < n; i++)
        if (first[slot_of[i]] == (uint32_t)i)
            unique_first[nu++] = i;
    return nu;
}

}  /* extern "C" */

#include "../../registrar_b200/csrc/regk_types.hpp"

extern "C" int emul_build_blob(const char *const *types, const uint32_t *lens, uint32_t ntypes, uint8_t *blob_out,
    uint32_t blob_cap, uint32_t *blob_len, uint32_t *max_escaped)
{
    std::vector<std::string> raw(ntypes);
    for (uint32_t i = 0; i < ntypes; i++)
        raw[i].assign(types[i], lens[i]);
    std::vector<uint8_t> blob;
    std::string why;
    int rc = regk::build_type_blob(raw, &blob, max_escaped, &why);
    if (rc)
        return rc;
    if (blob.size() > blob_cap)
        return 3;
    memcpy(blob_out, blob.data(), blob.size());
    *blob_len = (uint32_t)blob.size();
    return 0;
}

/* service-record payloads (regk_core.cuh emit_service): every record through the length sink, the byte sink and a
   word sink at output phase `phase`; returns 0 when the three agree, else 1 + the record index */
extern "C" uint64_t emul_services(uint64_t n, const uint8_t *srvce_bytes, const uint32_t *srvce_off, const uint8_t *proto_bytes,
    const uint32_t *proto_off, const uint32_t *port, const int32_t *ttl, const uint8_t *key_order, uint32_t phase,
    uint8_t *out_bytes, uint64_t *out_off)
{
    const GuardedWords ssrc{(const uint32_t *)srvce_bytes}, psrc{(const uint32_t *)proto_bytes};
    uint64_t base = 0;
    for (uint64_t r = 0; r < n; r++) {
        const uint32_t s0 = srvce_off[r], sl = srvce_off[r + 1] - s0, p0 = proto_off[r], pl = proto_off[r + 1] - p0;
        const uint32_t order = key_order ? key_order[r] : (uint32_t)KEY_ORDER_DEFAULT;
        if (!key_order_ok(order))
            return 1 + r;
        LenSink ls{0};
        emit_service(ssrc, s0, sl, psrc, p0, pl, port[r], ttl[r], order, ls, true);
        ByteSink bs;
        bs.init(out_bytes + base);
        emit_service(ssrc, s0, sl, psrc, p0, pl, port[r], ttl[r], order, bs, false);
        if ((uint64_t)(bs.p - (out_bytes + base)) != ls.n)
            return 1 + r;
        std::vector<uint32_t> img((phase + ls.n) / 4 + 8, 0);
        WordSink ws;
        ws.init(img.data(), phase);
        emit_service(ssrc, s0, sl, psrc, p0, pl, port[r], ttl[r], order, ws, false);
        ws.tail();
        if (memcmp((const uint8_t *)img.data() + phase, out_bytes + base, ls.n) != 0)
            return 1 + r;
        out_off[r] = base;
        base += ls.n;
    }
    out_off[n] = base;
    return 0;
}

/* the reader side (regk_decode_core.cuh) on the host, tile by tile as regk_decode_kernel runs it: out = n records of 10
   uint32 words, domains in slot layout (domain i at path_off[i]), ports in slot layout (element json_off[i] / 2).
   mode bit 0: host nodes; bit 1: the byte-wise route (global-memory fallback, guarded cursor) instead of the staged one */
extern "C" void emul_decode(uint64_t n, const uint8_t *path_bytes, const uint64_t *path_off, const uint8_t *json_bytes,
    const uint64_t *json_off, int mode, uint32_t *out, uint8_t *dom_bytes, uint32_t *ports)
{
    const bool host_nodes = mode & 1, bytewise = mode & 2;
    for (uint64_t r0 = 0; r0 < n; r0 += TILE) {
        const uint32_t nrec = (uint32_t)std::min<uint64_t>(TILE, n - r0);
        std::vector<Decoded> rec(nrec);
        for (auto &d : rec) {
            d.flags = 0;
            d.dom_len = d.host_pos = d.host_len = d.type_pos = d.type_len = d.addr_pos = d.addr_len = 0;
            d.ttl = INT32_MIN;
            d.nports = 0xFFFFFFFFu;
        }
        if (path_bytes && bytewise) {
            for (uint32_t t = 0; t < nrec; t++) {
                const uint64_t r = r0 + t;
                rec[t].flags |= decode_path(path_bytes + path_off[r], (uint32_t)(path_off[r + 1] - path_off[r]), host_nodes,
                    dom_bytes + path_off[r], rec[t]);
            }
        } else if (path_bytes) {
            const uint64_t P0 = path_off[r0], P1 = path_off[r0 + nrec];
            const uint32_t lead = (uint32_t)(P0 & 15u), np = (lead + (uint32_t)(P1 - P0) + 15u) & ~15u;
            std::vector<uint32_t> spath(np / 4 + 16, 0xA5A5A5A5u), sbits(np / 32 + 8, 0x3C3C3C3Cu), sdom(np / 4 + 16, 0u);
            memcpy((uint8_t *)spath.data() + lead, path_bytes + P0, (size_t)(P1 - P0));
            for (uint32_t tt = 0; tt < TILE; tt++)                   /* "threads" in a scrambled order */
                prepass_slashes(spath.data(), (uint16_t *)sbits.data(), np >> 4, (tt * 37u) % TILE, TILE);
            std::vector<WordSink> sinks(nrec);
            for (uint32_t t = nrec; t-- > 0;) {                      /* phase A in reverse order: neighbours must not clobber */
                const uint64_t r = r0 + t;
                const uint32_t off = lead + (uint32_t)(path_off[r] - P0);
                sinks[t].init(sdom.data(), off);
                rec[t].flags |= decode_path2(spath.data(), sbits.data(), off, (uint32_t)(path_off[r + 1] - path_off[r]), host_nodes,
                    sinks[t], rec[t]);
            }
            for (uint32_t t = 0; t < nrec; t++)                      /* phase B after the "barrier" */
                sinks[t].tail();
            memcpy(dom_bytes + P0, (const uint8_t *)sdom.data() + lead, (size_t)(P1 - P0));
        }
        for (uint32_t t = 0; t < nrec; t++) {
            const uint64_t r = r0 + t;
            if (json_bytes) {
                const uint8_t *j = json_bytes + json_off[r];
                const uint32_t jn = (uint32_t)(json_off[r + 1] - json_off[r]);
                rec[t].flags |= bytewise ? decode_payload<true>(j, jn, rec[t], ports + (json_off[r] >> 1))
                                         : decode_payload<false>(j, jn, rec[t], ports + (json_off[r] >> 1));
            }
            memcpy(out + 10 * r, &rec[t], sizeof(Decoded));
        }
    }
}
