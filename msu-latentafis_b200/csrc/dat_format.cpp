#include "dat_format.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <filesystem>

namespace lafis {
namespace {

struct Cursor {
    const uint8_t* p;
    size_t n, off = 0;
    bool fail = false;
    // std::ifstream::read semantics: a short read sets failbit and every later read is a no-op
    bool take(void* dst, size_t bytes) {
        if (fail || off + bytes > n) {
            fail = true;
            return false;
        }
        if (dst) std::memcpy(dst, p + off, bytes);
        off += bytes;
        return true;
    }
};

bool slurp(const std::string& path, std::vector<uint8_t>& buf) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long len = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.resize(len > 0 ? (size_t)len : 0);
    size_t got = len > 0 ? std::fread(buf.data(), 1, (size_t)len, f) : 0;
    std::fclose(f);
    return got == buf.size();
}

// one "n, x[n], y[n], ori[n], des_len, payload" record; `payload_elem` is 4 (float) or 1 (code byte).
// Returns false when the stream ran dry inside the record.
bool read_points(Cursor& c, int n, PointSet* keep, int16_t& des_len, int payload_elem) {
    std::vector<int16_t> x(n), y(n);
    std::vector<float> ori(n);
    c.take(x.data(), 2 * (size_t)n);
    c.take(y.data(), 2 * (size_t)n);
    c.take(ori.data(), 4 * (size_t)n);
    des_len = 0;
    c.take(&des_len, 2);
    if (c.fail || des_len <= 0) return false;
    const size_t bytes = (size_t)n * (size_t)des_len * (size_t)payload_elem;
    if (c.off + bytes > c.n) {
        c.fail = true;
        return false;
    }
    if (keep) {
        keep->x.swap(x);
        keep->y.swap(y);
        keep->ori.swap(ori);
        if (payload_elem == 4) {
            keep->des.resize((size_t)n * des_len);
            std::memcpy(keep->des.data(), c.p + c.off, bytes);
        } else {
            keep->codes.assign(c.p + c.off, c.p + c.off + bytes);
        }
    }
    c.off += bytes;
    return true;
}

}  // namespace

int read_rolled_dat(const std::string& path, RolledTemplate& out) {
    out = RolledTemplate();
    std::vector<uint8_t> buf;
    if (!slurp(path, buf)) {
        out.status = -1;
        return -3;
    }
    if (buf.size() <= 10) {  // matcher.cpp:899-902
        out.status = 1;
        return 1;
    }
    Cursor c{buf.data(), buf.size()};
    int16_t hdr[16];
    c.take(hdr, 2 * 16);  // 12-word header + h, w, blkH, blkW
    uint8_t n_minu_t = 0, n_tex_t = 0;
    if (!c.take(&n_minu_t, 1)) {  // header-only "empty" file (descriptor_PQ.py:190-193)
        out.status = 1;
        return 1;
    }
    for (int i = 0; i < n_minu_t; ++i) {
        int16_t n = 0;
        if (!c.take(&n, 2)) break;
        if (n <= 0) continue;  // matcher.cpp:936-937: record ends here, template not added
        if (n > kMaxMinutiae) {  // :938-942 returns 2 keeping what was added so far
            out.status = out.n_minu_templates ? 2 : 1;
            return 2;
        }
        int16_t dl = 0;
        const bool first = out.n_minu_templates == 0;
        if (!read_points(c, n, first ? &out.minu : nullptr, dl, 4)) {
            out = RolledTemplate();
            out.status = -1;  // truncated file: undefined in the reference, rejected here
            return -1;
        }
        if (first && dl != kDesLen) {
            out = RolledTemplate();
            out.status = -1;  // the reference asserts des_len equality at matcher.cpp:433
            return -1;
        }
        out.n_minu_templates++;
    }
    if (c.take(&n_tex_t, 1)) {
        for (int i = 0; i < n_tex_t; ++i) {
            int16_t n = 0;
            if (!c.take(&n, 2)) break;
            if (n <= 0) continue;
            if (n > kMaxMinutiae) {  // :966-970 returns -1 and the drivers empty the template
                out = RolledTemplate();
                out.status = -1;
                return -1;
            }
            int16_t dl = 0;
            if (!read_points(c, n, &out.tex, dl, 1) || dl != kSubs) {
                out = RolledTemplate();
                out.status = -1;
                return -1;
            }
            out.n_tex_templates = 1;
            // The reference over-reads 4*n*des_len bytes here (:975) and lands at EOF, so no
            // further texture record can be parsed; only template 0 is ever used (:413).
            break;
        }
    }
    if (out.tex.n() > kMaxTexture) {  // matcher.cpp:546-547: only the first 1000 points are used
        out.tex.x.resize(kMaxTexture);
        out.tex.y.resize(kMaxTexture);
        out.tex.ori.resize(kMaxTexture);
        out.tex.codes.resize((size_t)kMaxTexture * kSubs);
    }
    out.status = (out.n_minu_templates || out.n_tex_templates) ? 0 : 1;
    return 0;
}

int read_latent_dat(const std::string& path, LatentTemplate& out) {
    out = LatentTemplate();
    std::vector<uint8_t> buf;
    if (!slurp(path, buf)) {
        out.load_rc = -3;
        return -3;
    }
    if (buf.empty()) {  // matcher.cpp:798-801
        out.load_rc = 1;
        return 1;
    }
    Cursor c{buf.data(), buf.size()};
    int16_t hdr[16];
    c.take(hdr, 2 * 16);
    uint8_t n_minu_t = 0, n_tex_t = 0;
    if (!c.take(&n_minu_t, 1)) return 0;
    for (int i = 0; i < n_minu_t; ++i) {
        int16_t n = 0;
        if (!c.take(&n, 2)) break;
        if (n <= 0) continue;  // :835-836 skipped WITHOUT a slot: later templates shift down
        if (n > kMaxMinutiae) {
            out.load_rc = 2;
            return 2;  // the drivers ignore this return code (:150, :259)
        }
        PointSet* keep = nullptr;
        for (int s = 0; s < 3; ++s)
            if (out.n_minu_templates == kSelected[s]) keep = &out.minu[s];
        int16_t dl = 0;
        if (!read_points(c, n, keep, dl, 4) || (keep && dl != kDesLen)) {
            out = LatentTemplate();
            out.load_rc = -1;
            return -1;
        }
        out.n_minu_templates++;
    }
    if (c.take(&n_tex_t, 1)) {
        for (int i = 0; i < n_tex_t; ++i) {
            int16_t n = 0;
            if (!c.take(&n, 2)) break;
            if (n <= 0) continue;
            if (n > kMaxMinutiae) {
                out.load_rc = -1;
                return -1;
            }
            int16_t dl = 0;
            const bool first = out.n_tex_templates == 0;
            if (!read_points(c, n, first ? &out.tex : nullptr, dl, 4) || (first && dl != kDesLen)) {
                out = LatentTemplate();
                out.load_rc = -1;
                return -1;
            }
            out.n_tex_templates++;
        }
    }
    if (out.tex.n() > kMaxTexture) {  // matcher.cpp:544-545
        out.tex.x.resize(kMaxTexture);
        out.tex.y.resize(kMaxTexture);
        out.tex.ori.resize(kMaxTexture);
        out.tex.des.resize((size_t)kMaxTexture * kDesLen);
    }
    return 0;
}

int write_rolled_dat(const std::string& path, int h, int w, int blkH, int blkW, const PointSet& minu, const PointSet& tex) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return -3;
    bool ok = true;
    auto put = [&](const void* p, size_t bytes) {
        if (bytes) ok = ok && std::fwrite(p, 1, bytes, f) == bytes;
    };
    auto put_u16 = [&](unsigned v) {
        const uint16_t x = (uint16_t)v;
        put(&x, 2);
    };
    auto put_u8 = [&](unsigned v) {
        const uint8_t x = (uint8_t)v;
        put(&x, 1);
    };
    uint16_t header[12] = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // descriptor_PQ.py:186-189
    put(header, sizeof header);
    put_u16((unsigned)h);
    put_u16((unsigned)w);
    put_u16((unsigned)std::min(blkH, 50));
    put_u16((unsigned)std::min(blkW, 50));
    put_u8(1);  // one minutiae template
    const int nm = std::min(minu.n(), kMaxMinutiae);
    put_u16((unsigned)nm);
    if (nm > 0) {
        put(minu.x.data(), 2 * (size_t)nm);
        put(minu.y.data(), 2 * (size_t)nm);
        put(minu.ori.data(), 4 * (size_t)nm);
        put_u16(kDesLen);
        put(minu.des.data(), 4 * (size_t)nm * kDesLen);
    }
    put_u8(1);  // one texture template
    const int nt = std::min(tex.n(), kMaxMinutiae);
    put_u16((unsigned)nt);
    if (nt > 0) {
        put(tex.x.data(), 2 * (size_t)nt);
        put(tex.y.data(), 2 * (size_t)nt);
        put(tex.ori.data(), 4 * (size_t)nt);
        put_u16(kSubs);
        put(tex.codes.data(), (size_t)nt * kSubs);
    }
    ok = (std::fclose(f) == 0) && ok;
    return ok ? 0 : -3;
}

int write_latent_dat(const std::string& path, int h, int w, int blkH, int blkW, const std::vector<PointSet>& minu,
                     const std::vector<PointSet>& tex) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return -3;
    bool ok = true;
    auto put = [&](const void* p, size_t bytes) {
        if (bytes) ok = ok && std::fwrite(p, 1, bytes, f) == bytes;
    };
    auto put_u16 = [&](unsigned v) {
        const uint16_t x = (uint16_t)v;
        put(&x, 2);
    };
    auto put_u8 = [&](unsigned v) {
        const uint8_t x = (uint8_t)v;
        put(&x, 1);
    };
    uint16_t header[12] = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // descriptor_PQ.py:87-91
    put(header, sizeof header);
    if (minu.empty()) {  // :92-95 "empty" file
        uint16_t zeros[4] = {0, 0, 0, 0};
        put(zeros, sizeof zeros);
        ok = (std::fclose(f) == 0) && ok;
        return ok ? 0 : -3;
    }
    put_u16((unsigned)h);
    put_u16((unsigned)w);
    put_u16((unsigned)std::min(blkH, 50));
    put_u16((unsigned)std::min(blkW, 50));
    auto put_set = [&](const PointSet& s) {
        const int n = std::min(s.n(), kMaxMinutiae);
        put_u16((unsigned)n);
        if (n <= 0) return;
        put(s.x.data(), 2 * (size_t)n);
        put(s.y.data(), 2 * (size_t)n);
        put(s.ori.data(), 4 * (size_t)n);
        put_u16(kDesLen);
        put(s.des.data(), 4 * (size_t)n * kDesLen);
    };
    put_u8((unsigned)minu.size());
    for (const PointSet& s : minu) put_set(s);
    put_u8((unsigned)tex.size());
    for (const PointSet& s : tex) put_set(s);
    ok = (std::fclose(f) == 0) && ok;
    return ok ? 0 : -3;
}

int read_codebook(const std::string& path, std::vector<float>& cw, int& subs, int& clusters, int& sub_dim) {
    std::vector<uint8_t> buf;
    if (!slurp(path, buf)) return -3;
    if (buf.size() < 6) return -4;
    uint16_t h[3];
    std::memcpy(h, buf.data(), 6);
    subs = h[0];
    clusters = h[1];
    sub_dim = h[2];
    const size_t len = (size_t)subs * clusters * sub_dim;
    if (len == 0 || buf.size() < 6 + 4 * len) return -4;
    cw.resize(len);
    std::memcpy(cw.data(), buf.data() + 6, 4 * len);
    return 0;
}

std::string quoted_path(const std::string& p) {
    std::string out;
    out.reserve(p.size() + 2);
    out.push_back('"');
    for (char ch : p) {
        if (ch == '"' || ch == '&') out.push_back('&');
        out.push_back(ch);
    }
    out.push_back('"');
    return out;
}

void append_score_row(std::string& out, const std::string& quoted, float score) {
    char num[64];
    const int n = std::snprintf(num, sizeof num, "%.3f", (double)score);
    out.append(quoted);
    out.push_back(',');
    out.append(num, (size_t)(n > 0 ? n : 0));
    out.push_back('\n');
}

std::vector<std::string> list_dat_files(const std::string& dir) {
    std::vector<std::string> out;
    std::error_code ec;
    for (std::filesystem::directory_iterator it(dir, ec), end; !ec && it != end; it.increment(ec))
        if (it->path().extension() == ".dat") out.push_back(it->path().string());
    return out;
}

}  // namespace lafis
