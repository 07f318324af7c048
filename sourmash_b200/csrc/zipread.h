// zipread.h -- read-only access to ZIP containers of signatures (host code).
//
// The reference keeps signature collections in .zip files: members `signatures/<md5>.sig.gz`
// plus `SOURMASH-MANIFEST.csv` (src/sourmash/index/__init__.py:529-733 ZipFileLinearIndex,
// src/sourmash/sbt_storage.py:93-200 ZipStorage over src/core/src/storage/mod.rs:314-448, which
// maps the file and reads the central directory with the `piz` crate).  This is the same job on
// zlib only: map the file, walk the central directory (zip64 included), inflate members on any
// thread.  Entries keep the order of the central directory (ZipStorage._filenames / infolist()).
#pragma once
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

namespace smb {

// inflate [p, p + n) into out (appended).  window_bits: -15 raw deflate (zip members),
// 15 + 32 zlib / gzip with automatic header detection (multi-member gzip files are followed).
inline std::string inflate_all(const uint8_t* p, size_t n, int window_bits, size_t size_hint, std::string& out) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, window_bits) != Z_OK) return "zlib inflateInit2 failed";
    const size_t base = out.size();
    size_t cap = base + (size_hint ? size_hint : n * 4 + (1u << 16));
    out.resize(cap);
    size_t w = base, fed = 0;                            // zlib counts in 32 bits: feed and drain in chunks
    std::string err;
    for (;;) {
        if (zs.avail_in == 0 && fed < n) {
            const size_t c = std::min<size_t>(n - fed, 1u << 30);
            zs.next_in = (Bytef*)(p + fed); zs.avail_in = (uInt)c; fed += c;
        }
        if (w == cap) { cap = cap * 2 + (1u << 16); out.resize(cap); }
        const size_t room = std::min<size_t>(cap - w, 1u << 30);
        zs.next_out = (Bytef*)&out[w]; zs.avail_out = (uInt)room;
        const int r = inflate(&zs, Z_NO_FLUSH);
        w += room - zs.avail_out;
        if (r == Z_STREAM_END) {
            // gzip: concatenated members are one file (RFC 1952)
            const size_t at = fed - zs.avail_in;
            if (window_bits > 15 && n - at >= 2 && p[at] == 0x1f && p[at + 1] == 0x8b) {
                if (inflateReset(&zs) != Z_OK) { err = "zlib inflateReset failed"; break; }
                continue;
            }
            break;
        }
        if (r == Z_OK) continue;
        if (r == Z_BUF_ERROR && (zs.avail_out == 0 || (zs.avail_in == 0 && fed < n))) continue;
        err = "compressed stream is corrupt or truncated";
        break;
    }
    inflateEnd(&zs);
    out.resize(err.empty() ? w : base);
    return err;
}

struct ZipMember {
    std::string name;
    uint16_t method = 0;                 // 0 stored, 8 deflate
    uint32_t crc = 0;
    uint64_t csize = 0, usize = 0, local_off = 0;
    bool is_dir() const { return !name.empty() && name.back() == '/'; }
};

class ZipArchive {
  public:
    std::vector<ZipMember> members;                       // central-directory order
    std::unordered_map<std::string, size_t> by_name;      // a repeated name resolves to its last entry

    ZipArchive() {}
    ZipArchive(const ZipArchive&) = delete;
    ZipArchive& operator=(const ZipArchive&) = delete;
    ~ZipArchive() { if (map_) munmap(map_, len_); }

    static bool has_magic(const uint8_t* p, size_t n) {   // local header, or the end record of an empty archive
        return n >= 4 && p[0] == 'P' && p[1] == 'K' && ((p[2] == 3 && p[3] == 4) || (p[2] == 5 && p[3] == 6));
    }

    std::string open(const char* path) {
        int fd = ::open(path, O_RDONLY);
        if (fd < 0) return std::string("cannot open ") + path;
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size < 22) { close(fd); return std::string(path) + ": not a zip file"; }
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        if (m == MAP_FAILED) return std::string("cannot map ") + path;
        map_ = m; len_ = (size_t)st.st_size;
        std::string e = open_memory((const uint8_t*)m, len_);
        return e.empty() ? e : std::string(path) + ": " + e;
    }

    // the caller keeps [p, p + n) alive
    std::string open_memory(const uint8_t* p, size_t n) {
        p_ = p; n_ = n;
        if (n < 22) return "not a zip file";
        // end of central directory: last 22 bytes + up to 64 KiB of archive comment
        size_t eocd = n;
        const size_t lowest = n - 22 > 65535 ? n - 22 - 65535 : 0;
        for (size_t at = n - 22 + 1; at-- > lowest;) {
            if (rd32(p + at) == 0x06054b50u && at + 22 + rd16(p + at + 20) <= n) { eocd = at; break; }
        }
        if (eocd == n) return "not a zip file (no end-of-central-directory record)";
        uint64_t n_entries = rd16(p + eocd + 10), cd_size = rd32(p + eocd + 12), cd_off = rd32(p + eocd + 16);
        if (eocd >= 20 && rd32(p + eocd - 20) == 0x07064b50u) {                  // zip64 locator
            const uint64_t at64 = rd64(p + eocd - 20 + 8);
            if (at64 > n || n - at64 < 56 || rd32(p + at64) != 0x06064b50u) return "bad zip64 end-of-central-directory record";
            n_entries = rd64(p + at64 + 32); cd_size = rd64(p + at64 + 40); cd_off = rd64(p + at64 + 48);
        }
        if (cd_off > n || cd_size > n - cd_off) return "central directory lies outside the file";
        members.clear(); by_name.clear();
        members.reserve((size_t)std::min<uint64_t>(n_entries, 1u << 24));
        const uint8_t* q = p + cd_off;
        const uint8_t* end = q + cd_size;
        for (uint64_t i = 0; i < n_entries; ++i) {
            if (end - q < 46 || rd32(q) != 0x02014b50u) return "bad central directory entry";
            ZipMember m;
            m.method = rd16(q + 10); m.crc = rd32(q + 16);
            m.csize = rd32(q + 20); m.usize = rd32(q + 24);
            const size_t nlen = rd16(q + 28), elen = rd16(q + 30), clen = rd16(q + 32);
            m.local_off = rd32(q + 42);
            if ((size_t)(end - q) < 46 + nlen + elen + clen) return "bad central directory entry";
            m.name.assign((const char*)q + 46, nlen);
            // zip64 extended information: 8-byte fields for every 32-bit field that is saturated
            const uint8_t* x = q + 46 + nlen;
            const uint8_t* xe = x + elen;
            while (xe - x >= 4) {
                const uint16_t id = rd16(x), sz = rd16(x + 2);
                const uint8_t* f = x + 4;
                if (f + sz > xe) break;
                if (id == 0x0001) {
                    const uint8_t* fe = f + sz;
                    if (m.usize == 0xFFFFFFFFu && fe - f >= 8) { m.usize = rd64(f); f += 8; }
                    if (m.csize == 0xFFFFFFFFu && fe - f >= 8) { m.csize = rd64(f); f += 8; }
                    if (m.local_off == 0xFFFFFFFFu && fe - f >= 8) { m.local_off = rd64(f); f += 8; }
                }
                x += 4 + sz;
            }
            q += 46 + nlen + elen + clen;
            by_name[m.name] = members.size();
            members.push_back(std::move(m));
        }
        return "";
    }

    const ZipMember* find(const std::string& name) const {
        auto it = by_name.find(name);
        return it == by_name.end() ? nullptr : &members[it->second];
    }

    // member bytes appended to out; checks the CRC like zipfile / piz do
    std::string read(const ZipMember& m, std::string& out) const {
        if (m.local_off > n_ || n_ - m.local_off < 30 || rd32(p_ + m.local_off) != 0x04034b50u)
            return m.name + ": bad local file header";
        const uint8_t* h = p_ + m.local_off;
        const uint64_t data = m.local_off + 30 + rd16(h + 26) + rd16(h + 28);
        if (data > n_ || m.csize > n_ - data) return m.name + ": member data lies outside the file";
        const size_t base = out.size();
        if (m.method == 0) {
            out.append((const char*)p_ + data, (size_t)m.csize);
        } else if (m.method == 8) {
            // the declared size is only a hint for the first allocation: deflate expands at most ~1032x, so a
            // crafted header cannot ask for more than the compressed bytes could possibly produce
            const uint64_t most = m.csize * 1032 + (1u << 16);
            std::string e = inflate_all(p_ + data, (size_t)m.csize, -15, (size_t)std::min<uint64_t>(m.usize, most) + 16, out);
            if (!e.empty()) return m.name + ": " + e;
        } else {
            return m.name + ": unsupported compression method " + std::to_string(m.method);
        }
        uLong crc = crc32(0L, Z_NULL, 0);
        for (size_t at = base; at < out.size();) {
            const size_t c = std::min<size_t>(out.size() - at, 1u << 30);
            crc = crc32(crc, (const Bytef*)out.data() + at, (uInt)c);
            at += c;
        }
        if ((uint32_t)crc != m.crc) { out.resize(base); return m.name + ": bad CRC-32"; }
        return "";
    }

  private:
    static uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
    static uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
    static uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

    const uint8_t* p_ = nullptr;
    size_t n_ = 0;
    void* map_ = nullptr;
    size_t len_ = 0;
};

// Rows of a manifest (src/sourmash/manifest.py:90-140: a "# SOURMASH-MANIFEST-VERSION: x" line,
// a CSV header, one row per sketch): the distinct internal_location values in row order
// (manifest.py:365-374) and the md5 of every row (manifest.py:376-379).
struct ManifestIndex {
    std::vector<std::string> locations;
    std::vector<std::string> md5s;
};

inline std::string parse_manifest_csv(const std::string& text, ManifestIndex& out) {
    // RFC 4180 fields: quoted fields may hold commas, doubled quotes and line breaks
    std::vector<std::vector<std::string>> rows;
    std::vector<std::string> cur;
    std::string field;
    bool quoted = false, any = false;
    auto end_row = [&] {
        cur.push_back(std::move(field)); field.clear();
        rows.push_back(std::move(cur)); cur.clear();
        any = false;
    };
    for (size_t i = 0; i < text.size(); ++i) {
        const char c = text[i];
        if (quoted) {
            if (c == '"') {
                if (i + 1 < text.size() && text[i + 1] == '"') { field.push_back('"'); ++i; }
                else quoted = false;
            } else field.push_back(c);
            continue;
        }
        if (c == '"') { quoted = true; any = true; }
        else if (c == ',') { cur.push_back(std::move(field)); field.clear(); any = true; }
        else if (c == '\n') { if (any || !field.empty() || !cur.empty()) end_row(); }
        else if (c == '\r') {}
        else { field.push_back(c); any = true; }
    }
    if (any || !field.empty() || !cur.empty()) end_row();
    size_t r = 0;
    if (r < rows.size() && !rows[r].empty() && !rows[r][0].empty() && rows[r][0][0] == '#') ++r;   // version line
    if (r >= rows.size()) return "manifest has no header row";
    int loc_col = -1, md5_col = -1;
    for (size_t c = 0; c < rows[r].size(); ++c) {
        if (rows[r][c] == "internal_location") loc_col = (int)c;
        else if (rows[r][c] == "md5") md5_col = (int)c;
    }
    if (loc_col < 0 || md5_col < 0) return "manifest lacks the internal_location / md5 columns";
    std::unordered_map<std::string, bool> seen;
    for (++r; r < rows.size(); ++r) {
        const auto& row = rows[r];
        if ((int)row.size() <= std::max(loc_col, md5_col)) continue;
        if (seen.emplace(row[loc_col], true).second) out.locations.push_back(row[loc_col]);
        out.md5s.push_back(row[md5_col]);
    }
    return "";
}

}  // namespace smb
