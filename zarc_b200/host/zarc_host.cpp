// See zarc_host.hpp.  Format code follows SURVEY.md App. A (bytes as the reference implements them).
#include "zarc_host.hpp"

#include <algorithm>
#include <cerrno>
#include <cstring>
#include <ctime>
#include <fcntl.h>
#include <grp.h>
#include <pwd.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "zarcgpu_dl.hpp"

namespace zarc {

const uint8_t FILE_MAGIC[12] = {0x50, 0x2A, 0x4D, 0x18, 0x04, 0x00, 0x00, 0x00, 0x65, 0xAA, 0xDC, ZARC_VERSION};

// ================================================================================================
// CBOR (RFC 8949), the subset minicbor emits for these structs; the reader also takes indefinite
// lengths and null for absent fields (SURVEY.md App. A caveat i)
namespace cbor {
struct Writer {
	Bytes out;
	void head(uint8_t major, uint64_t v) {
		uint8_t m = (uint8_t)(major << 5);
		if (v < 24) out.push_back(m | (uint8_t)v);
		else if (v <= 0xff) { out.push_back(m | 24); out.push_back((uint8_t)v); }
		else if (v <= 0xffff) { out.push_back(m | 25); out.push_back((uint8_t)(v >> 8)); out.push_back((uint8_t)v); }
		else if (v <= 0xffffffffull) { out.push_back(m | 26); for (int s = 24; s >= 0; s -= 8) out.push_back((uint8_t)(v >> s)); }
		else { out.push_back(m | 27); for (int s = 56; s >= 0; s -= 8) out.push_back((uint8_t)(v >> s)); }
	}
	void uint(uint64_t v) { head(0, v); }
	void bytes(const uint8_t* p, size_t n) { head(2, n); out.insert(out.end(), p, p + n); }
	void text(const std::string& s) { head(3, s.size()); out.insert(out.end(), s.begin(), s.end()); }
	void array(uint64_t n) { head(4, n); }
	void map(uint64_t n) { head(5, n); }
	void tag(uint64_t t) { head(6, t); }
	void boolean(bool b) { out.push_back(b ? 0xf5 : 0xf4); }
	void null() { out.push_back(0xf6); }
};

struct Reader {
	const uint8_t* p;
	const uint8_t* end;
	[[noreturn]] static void fail(const char* what) { throw Error(std::string("cbor: ") + what); }
	uint8_t peek() const {
		if (p >= end) fail("unexpected end of input");
		return *p;
	}
	uint8_t major() const { return peek() >> 5; }
	bool is_null() const { return peek() == 0xf6 || peek() == 0xf7; }
	bool is_break() const { return peek() == 0xff; }
	// reads a head; returns false for an indefinite length
	bool head(uint8_t& major_out, uint64_t& v) {
		uint8_t b = peek();
		p++;
		major_out = b >> 5;
		uint8_t ai = b & 31;
		if (ai < 24) { v = ai; return true; }
		if (ai == 31) { v = 0; return false; }
		if (ai > 27) fail("reserved additional information");
		size_t n = (size_t)1 << (ai - 24);
		if ((size_t)(end - p) < n) fail("unexpected end of input");
		v = 0;
		for (size_t i = 0; i < n; i++) v = (v << 8) | *p++;
		return true;
	}
	uint64_t uint() {
		uint8_t m; uint64_t v;
		if (!head(m, v) || m != 0) fail("expected an unsigned integer");
		return v;
	}
	int64_t integer() {
		uint8_t m; uint64_t v;
		if (!head(m, v) || m > 1) fail("expected an integer");
		return m == 0 ? (int64_t)v : -1 - (int64_t)v;
	}
	std::string string_like(uint8_t want) {  // definite or indefinite byte/text string
		uint8_t m; uint64_t v;
		bool def = head(m, v);
		if (m != want) fail("expected a string");
		std::string s;
		if (def) {
			if ((uint64_t)(end - p) < v) fail("string runs past the input");
			s.assign((const char*)p, (size_t)v);
			p += v;
		} else {
			while (!is_break()) s += string_like(want);
			p++;
		}
		return s;
	}
	uint64_t tag() {
		uint8_t m; uint64_t v;
		if (!head(m, v) || m != 6) fail("expected a tag");
		return v;
	}
	bool boolean() {
		uint8_t b = peek();
		if (b != 0xf4 && b != 0xf5) fail("expected a boolean");
		p++;
		return b == 0xf5;
	}
	// container headers: the count, or UINT64_MAX for indefinite (then test is_break())
	uint64_t container(uint8_t want) {
		uint8_t m; uint64_t v;
		bool def = head(m, v);
		if (m != want) fail(want == 4 ? "expected an array" : "expected a map");
		return def ? v : UINT64_MAX;
	}
	bool more(uint64_t& left) {  // iteration helper for both container flavours
		if (left == UINT64_MAX) {
			if (is_break()) { p++; return false; }
			return true;
		}
		if (left == 0) return false;
		left--;
		return true;
	}
	void skip() {
		uint8_t b = peek();
		uint8_t m = b >> 5;
		if (m == 7) {
			uint8_t ai = b & 31;
			p++;
			size_t n = ai == 24 ? 1 : ai == 25 ? 2 : ai == 26 ? 4 : ai == 27 ? 8 : 0;
			if ((size_t)(end - p) < n) fail("unexpected end of input");
			p += n;
			return;
		}
		if (m == 2 || m == 3) { string_like(m); return; }
		uint8_t mm; uint64_t v;
		bool def = head(mm, v);
		if (m == 0 || m == 1) return;
		if (m == 6) { skip(); return; }
		uint64_t items = m == 5 ? 2 : 1;
		if (def) for (uint64_t i = 0; i < v * items; i++) skip();
		else { while (!is_break()) skip(); p++; }
	}
};
}  // namespace cbor

// ================================================================================================
// integrity.rs
bool Digest::operator==(const Digest& o) const {
	if (bytes.size() != o.bytes.size()) return false;
	uint8_t acc = 0;
	for (size_t i = 0; i < bytes.size(); i++) acc |= bytes[i] ^ o.bytes[i];
	return acc == 0;
}
size_t DigestHash::operator()(const Digest& d) const {
	size_t h = 0;
	memcpy(&h, d.bytes.data(), std::min(sizeof h, d.bytes.size()));
	return h;
}
static const char B64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
std::string Digest::base64() const {  // base64ct::Base64 (standard alphabet, padded) as the CLI prints it
	std::string s;
	size_t i = 0;
	for (; i + 2 < bytes.size(); i += 3) {
		uint32_t v = bytes[i] << 16 | bytes[i + 1] << 8 | bytes[i + 2];
		s += B64[v >> 18]; s += B64[(v >> 12) & 63]; s += B64[(v >> 6) & 63]; s += B64[v & 63];
	}
	if (i + 1 == bytes.size()) {
		uint32_t v = bytes[i] << 16;
		s += B64[v >> 18]; s += B64[(v >> 12) & 63]; s += "==";
	} else if (i + 2 == bytes.size()) {
		uint32_t v = bytes[i] << 16 | bytes[i + 1] << 8;
		s += B64[v >> 18]; s += B64[(v >> 12) & 63]; s += B64[(v >> 6) & 63]; s += '=';
	}
	return s;
}
Digest Digest::from_base64(const std::string& s) {
	Digest d;
	uint32_t acc = 0;
	int bits = 0;
	for (char c : s) {
		if (c == '=') break;
		const char* q = strchr(B64, c);
		if (!q || !c) throw Error("invalid base64 digest");
		acc = (acc << 6) | (uint32_t)(q - B64);
		bits += 6;
		if (bits >= 8) {
			bits -= 8;
			d.bytes.push_back((uint8_t)(acc >> bits));
		}
	}
	return d;
}

// ================================================================================================
// strings.rs
static bool valid_utf8(const std::string& s) {
	const uint8_t* p = (const uint8_t*)s.data();
	size_t n = s.size(), i = 0;
	while (i < n) {
		uint8_t c = p[i];
		size_t len = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 0;
		if (!len || i + len > n) return false;
		uint32_t cp = len == 1 ? c : c & (0xff >> (len + 1));
		for (size_t k = 1; k < len; k++) {
			if ((p[i + k] >> 6) != 2) return false;
			cp = (cp << 6) | (p[i + k] & 63);
		}
		if ((len == 2 && cp < 0x80) || (len == 3 && cp < 0x800) || (len == 4 && cp < 0x10000) || cp > 0x10ffff || (cp >= 0xd800 && cp <= 0xdfff))
			return false;
		i += len;
	}
	return true;
}
CborString CborString::from_maybe_utf8(const std::string& raw) {
	CborString c;
	c.is_text = valid_utf8(raw);
	c.data = raw;
	return c;
}
bool CborString::operator<(const CborString& o) const {  // derive(Ord): Text < Binary, then contents
	if (is_text != o.is_text) return is_text;
	return data < o.data;
}
Pathname Pathname::from_normal_components(const std::string& path) {
	Pathname p;
	size_t i = 0;
	while (i <= path.size()) {
		size_t j = path.find('/', i);
		if (j == std::string::npos) j = path.size();
		std::string comp = path.substr(i, j - i);
		if (!comp.empty() && comp != "." && comp != "..") p.components.push_back(CborString::from_maybe_utf8(comp));
		i = j + 1;
	}
	return p;
}
std::string Pathname::to_path() const {
	std::string s;
	for (const auto& c : components) {
		if (!c.data.empty() && c.data[0] == '/') s = c.data;  // PathBuf::push of an absolute component replaces
		else {
			if (!s.empty() && s.back() != '/') s += '/';
			s += c.data;
		}
	}
	return s;
}

// The extraction path of an entry read from an (untrusted) archive: "." and empty components are dropped, and a
// component that is "..", holds a '/' or a NUL byte makes the whole name unsafe (nullopt).  Deliberate divergence
// from the reference, whose unpack.rs:60-62 hands Pathname::to_path() to the filesystem verbatim (".." kept, an
// absolute component replaces the path): that is an arbitrary-file-overwrite primitive for a hostile .zarc.
std::optional<std::string> Pathname::to_safe_path() const {
	std::string s;
	for (const auto& c : components) {
		if (c.data.empty() || c.data == ".") continue;
		if (c.data == ".." || c.data.find('/') != std::string::npos || c.data.find('\0') != std::string::npos) return std::nullopt;
		if (!s.empty()) s += '/';
		s += c.data;
	}
	if (s.empty()) return std::nullopt;
	return s;
}

// ================================================================================================
// timestamps.rs
static int64_t days_from_civil(int64_t y, unsigned m, unsigned d) {
	y -= m <= 2;
	const int64_t era = (y >= 0 ? y : y - 399) / 400;
	const unsigned yoe = (unsigned)(y - era * 400);
	const unsigned doy = (153 * (m > 2 ? m - 3 : m + 9) + 2) / 5 + d - 1;
	const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
	return era * 146097 + (int64_t)doe - 719468;
}
static void civil_from_days(int64_t z, int64_t& y, unsigned& m, unsigned& d) {
	z += 719468;
	const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
	const unsigned doe = (unsigned)(z - era * 146097);
	const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
	y = (int64_t)yoe + era * 400;
	const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
	const unsigned mp = (5 * doy + 2) / 153;
	d = doy - (153 * mp + 2) / 5 + 1;
	m = mp < 10 ? mp + 3 : mp - 9;
	y += m <= 2;
}
Timestamp Timestamp::now() {
	timespec ts;
	clock_gettime(CLOCK_REALTIME, &ts);
	return Timestamp{(int64_t)ts.tv_sec, (uint32_t)ts.tv_nsec};
}
std::string Timestamp::to_rfc3339() const {  // chrono's to_rfc3339(): SecondsFormat::AutoSi, "+00:00"
	int64_t days = secs >= 0 ? secs / 86400 : -((-secs + 86399) / 86400);
	int64_t rem = secs - days * 86400;
	int64_t y; unsigned mo, d;
	civil_from_days(days, y, mo, d);
	char buf[80];
	int n = snprintf(buf, sizeof buf, "%04lld-%02u-%02uT%02d:%02d:%02d", (long long)y, mo, d, (int)(rem / 3600), (int)(rem / 60 % 60), (int)(rem % 60));
	std::string s(buf, (size_t)n);
	if (nanos) {
		if (nanos % 1000000 == 0) snprintf(buf, sizeof buf, ".%03u", nanos / 1000000);
		else if (nanos % 1000 == 0) snprintf(buf, sizeof buf, ".%06u", nanos / 1000);
		else snprintf(buf, sizeof buf, ".%09u", nanos);
		s += buf;
	}
	return s + "+00:00";
}
Timestamp Timestamp::parse_rfc3339(const std::string& s) {
	int y, mo, d, h, mi, se;
	int pos = 0;
	if (sscanf(s.c_str(), "%d-%d-%d%*1[Tt ]%d:%d:%d%n", &y, &mo, &d, &h, &mi, &se, &pos) != 6) throw Error("invalid RFC 3339 timestamp: " + s);
	uint32_t nanos = 0;
	size_t i = (size_t)pos;
	if (i < s.size() && s[i] == '.') {
		i++;
		uint32_t scale = 100000000;
		while (i < s.size() && isdigit((unsigned char)s[i])) {
			nanos += (uint32_t)(s[i] - '0') * scale;
			scale /= 10;
			i++;
		}
	}
	int64_t offset = 0;
	if (i < s.size() && (s[i] == '+' || s[i] == '-')) {
		int oh = 0, om = 0;
		if (sscanf(s.c_str() + i + 1, "%d:%d", &oh, &om) != 2) throw Error("invalid RFC 3339 offset: " + s);
		offset = (oh * 3600 + om * 60) * (s[i] == '-' ? -1 : 1);
	}
	int64_t secs = days_from_civil(y, (unsigned)mo, (unsigned)d) * 86400 + h * 3600 + mi * 60 + se - offset;
	return Timestamp{secs, nanos};
}

// ================================================================================================
// CBOR forms of the directory structs
static void put(cbor::Writer& w, const CborString& s) {
	if (s.is_text) w.text(s.data);
	else w.bytes((const uint8_t*)s.data.data(), s.data.size());
}
static void put(cbor::Writer& w, const Timestamp& t) {
	w.tag(0);
	w.text(t.to_rfc3339());
}
static void put(cbor::Writer& w, const AttributeMap& m) {
	w.map(m.size());
	for (const auto& kv : m) {
		w.text(kv.first);
		if (kv.second.is_bool) w.boolean(kv.second.b);
		else put(w, kv.second.s);
	}
}
static void put(cbor::Writer& w, const PosixOwner& o) {  // posix_owner.rs:225-247
	w.array((o.id ? 1 : 0) + (o.name ? 1 : 0));
	if (o.id) w.uint(*o.id);
	if (o.name) put(w, *o.name);
}
static void put(cbor::Writer& w, const Timestamps& t) {
	w.map((t.created ? 1 : 0) + (t.modified ? 1 : 0) + (t.accessed ? 1 : 0));
	if (t.created) { w.uint(1); put(w, *t.created); }
	if (t.modified) { w.uint(2); put(w, *t.modified); }
	if (t.accessed) { w.uint(3); put(w, *t.accessed); }
}
static void put(cbor::Writer& w, const SpecialFile& s) {  // #[cbor(array)]: positional, trailing None dropped
	size_t n = s.link_target ? 2 : s.kind ? 1 : 0;
	w.array(n);
	if (n >= 1) {
		if (s.kind) w.uint((uint8_t)*s.kind);
		else w.null();
	}
	if (n >= 2) {
		const LinkTarget& t = *s.link_target;
		if (t.is_components) {
			w.array(t.components.size());
			for (const auto& c : t.components) put(w, c);
		} else put(w, t.full_path);
	}
}

static CborString get_string(cbor::Reader& r) {
	CborString s;
	uint8_t m = r.major();
	if (m != 2 && m != 3) cbor::Reader::fail("expected a text or byte string");
	s.is_text = m == 3;
	s.data = r.string_like(m);
	return s;
}
static Timestamp get_timestamp(cbor::Reader& r) {  // timestamps.rs:80-125: tag 0 text, or tag 1 number
	uint64_t tag = r.tag();
	if (tag == 0) return Timestamp::parse_rfc3339(r.string_like(3));
	if (tag == 1) {
		if (r.major() <= 1) return Timestamp{r.integer(), 0};
		uint8_t b = r.peek();
		r.p++;
		if (b == 0xfa) {
			if (r.end - r.p < 4) cbor::Reader::fail("unexpected end of input");
			uint32_t u = 0;
			for (int i = 0; i < 4; i++) u = (u << 8) | *r.p++;
			float f;
			memcpy(&f, &u, 4);
			return Timestamp{(int64_t)f, (uint32_t)((f - (float)(int64_t)f) * 1.0e9f)};
		}
		if (b == 0xfb) {
			if (r.end - r.p < 8) cbor::Reader::fail("unexpected end of input");
			uint64_t u = 0;
			for (int i = 0; i < 8; i++) u = (u << 8) | *r.p++;
			double f;
			memcpy(&f, &u, 8);
			return Timestamp{(int64_t)f, (uint32_t)((f - (double)(int64_t)f) * 1.0e9)};
		}
		cbor::Reader::fail("unsupported epoch timestamp type");
	}
	cbor::Reader::fail("expected Timestamp or DateTime tag");
}
static AttributeMap get_attributes(cbor::Reader& r) {
	AttributeMap m;
	uint64_t left = r.container(5);
	while (r.more(left)) {
		std::string k = r.string_like(3);
		AttributeValue v;
		if (r.major() == 7) {
			v.is_bool = true;
			v.b = r.boolean();
		} else v.s = get_string(r);
		m.emplace(std::move(k), std::move(v));
	}
	return m;
}
static PosixOwner get_owner(cbor::Reader& r) {  // posix_owner.rs:249-278
	PosixOwner o;
	uint64_t left = r.container(4);
	while (r.more(left)) {
		if (r.major() == 0) o.id = r.uint();
		else if (r.major() == 3) o.name = get_string(r);
		else cbor::Reader::fail("unexpected type in a POSIX owner");
	}
	return o;
}
static Timestamps get_timestamps(cbor::Reader& r) {
	Timestamps t;
	uint64_t left = r.container(5);
	while (r.more(left)) {
		uint64_t k = r.uint();
		if (r.is_null()) { r.skip(); continue; }
		if (k == 1) t.created = get_timestamp(r);
		else if (k == 2) t.modified = get_timestamp(r);
		else if (k == 3) t.accessed = get_timestamp(r);
		else r.skip();
	}
	return t;
}
static SpecialFile get_special(cbor::Reader& r) {
	SpecialFile s;
	uint64_t left = r.container(4);
	size_t idx = 0;
	while (r.more(left)) {
		if (r.is_null()) r.skip();
		else if (idx == 0) s.kind = (SpecialFileKind)r.uint();
		else if (idx == 1) {
			LinkTarget t;
			if (r.major() == 4) {  // (the reference's decoder is todo!() here, specials.rs:193-196; the writer does emit it)
				t.is_components = true;
				uint64_t n = r.container(4);
				while (r.more(n)) t.components.push_back(get_string(r));
			} else t.full_path = get_string(r);
			s.link_target = std::move(t);
		} else r.skip();
		idx++;
	}
	return s;
}

static Bytes frame_element(ElementKind kind, const Bytes& payload) {  // elements.rs:10-25
	if (payload.size() > 0xffff) throw Error("directory element too large (out of range integral type conversion attempted)");
	Bytes b;
	b.reserve(payload.size() + 4);
	b.push_back((uint8_t)kind);
	b.push_back((uint8_t)payload.size());
	b.push_back((uint8_t)(payload.size() >> 8));
	b.push_back(0);
	b.insert(b.end(), payload.begin(), payload.end());
	return b;
}
Bytes encode_element(const Edition& e) {
	cbor::Writer w;
	w.map(3 + (e.user_metadata ? 1 : 0));
	w.uint(0); w.uint(e.number);
	w.uint(1); put(w, e.written_at);
	w.uint(2); w.uint((uint8_t)e.digest_type);
	if (e.user_metadata) { w.uint(10); put(w, *e.user_metadata); }
	return frame_element(ElementKind::Edition, w.out);
}
Bytes encode_element(const File& f) {
	cbor::Writer w;
	size_t n = 2 + (f.digest ? 1 : 0) + (f.mode ? 1 : 0) + (f.user ? 1 : 0) + (f.group ? 1 : 0) + (f.timestamps ? 1 : 0) + (f.special ? 1 : 0) +
	           (f.user_metadata ? 1 : 0) + (f.attributes ? 1 : 0) + (f.extended_attributes ? 1 : 0);
	w.map(n);
	w.uint(0); w.uint(f.edition);
	w.uint(1);
	w.array(f.name.components.size());
	for (const auto& c : f.name.components) put(w, c);
	if (f.digest) { w.uint(2); w.bytes(f.digest->bytes.data(), f.digest->bytes.size()); }
	if (f.mode) { w.uint(3); w.uint(*f.mode); }
	if (f.user) { w.uint(4); put(w, *f.user); }
	if (f.group) { w.uint(5); put(w, *f.group); }
	if (f.timestamps) { w.uint(6); put(w, *f.timestamps); }
	if (f.special) { w.uint(7); put(w, *f.special); }
	if (f.user_metadata) { w.uint(10); put(w, *f.user_metadata); }
	if (f.attributes) { w.uint(11); put(w, *f.attributes); }
	if (f.extended_attributes) { w.uint(12); put(w, *f.extended_attributes); }
	return frame_element(ElementKind::File, w.out);
}
Bytes encode_element(const Frame& f) {
	cbor::Writer w;
	w.map(5);
	w.uint(0); w.uint(f.edition);
	w.uint(1); w.uint(f.offset);
	w.uint(2); w.bytes(f.digest.bytes.data(), f.digest.bytes.size());
	w.uint(3); w.uint(f.length);
	w.uint(4); w.uint(f.uncompressed);
	return frame_element(ElementKind::Frame, w.out);
}

void parse_directory_stream(const uint8_t* p, size_t n, Directory& out) {
	size_t pos = 0;
	while (pos < n) {
		if (n - pos < 4) throw Error("parse error: truncated directory element header");
		uint8_t kind = p[pos];
		size_t len = p[pos + 1] | (size_t)p[pos + 2] << 8;
		pos += 4;
		if (n - pos < len) throw Error("parse error: directory element runs past the stream");
		cbor::Reader r{p + pos, p + pos + len};
		pos += len;
		if (kind == (uint8_t)ElementKind::Edition) {
			Edition e;
			uint64_t left = r.container(5);
			while (r.more(left)) {
				uint64_t k = r.uint();
				if (r.is_null()) { r.skip(); continue; }
				if (k == 0) e.number = (uint16_t)r.uint();
				else if (k == 1) e.written_at = get_timestamp(r);
				else if (k == 2) e.digest_type = (DigestType)r.uint();
				else if (k == 10) e.user_metadata = get_attributes(r);
				else r.skip();
			}
			out.editions[e.number] = std::move(e);
		} else if (kind == (uint8_t)ElementKind::Frame) {
			Frame f;
			uint64_t left = r.container(5);
			while (r.more(left)) {
				uint64_t k = r.uint();
				if (k == 0) f.edition = (uint16_t)r.uint();
				else if (k == 1) f.offset = r.uint();
				else if (k == 2) { std::string s = r.string_like(2); f.digest.bytes.assign(s.begin(), s.end()); }
				else if (k == 3) f.length = r.uint();
				else if (k == 4) f.uncompressed = r.uint();
				else r.skip();
			}
			// the digest is untrusted input: its length is fixed by the digest type (integrity.rs:98-107), and the
			// decoder copies exactly that many bytes out of it
			if (f.digest.bytes.size() != DIGEST_LEN) throw Error("parse error: frame digest has the wrong length");
			Digest key = f.digest;
			out.frames[key] = std::move(f);
		} else if (kind == (uint8_t)ElementKind::File) {
			File f;
			uint64_t left = r.container(5);
			while (r.more(left)) {
				uint64_t k = r.uint();
				if (r.is_null()) { r.skip(); continue; }
				if (k == 0) f.edition = (uint16_t)r.uint();
				else if (k == 1) {
					uint64_t m = r.container(4);
					while (r.more(m)) f.name.components.push_back(get_string(r));
				} else if (k == 2) { std::string s = r.string_like(2); Digest d; d.bytes.assign(s.begin(), s.end()); f.digest = std::move(d); }
				else if (k == 3) f.mode = (uint32_t)r.uint();
				else if (k == 4) f.user = get_owner(r);
				else if (k == 5) f.group = get_owner(r);
				else if (k == 6) f.timestamps = get_timestamps(r);
				else if (k == 7) f.special = get_special(r);
				else if (k == 10) f.user_metadata = get_attributes(r);
				else if (k == 11) f.attributes = get_attributes(r);
				else if (k == 12) f.extended_attributes = get_attributes(r);
				else r.skip();
			}
			if (f.digest && f.digest->bytes.size() != DIGEST_LEN) throw Error("parse error: file digest has the wrong length");
			size_t index = out.files.size();
			out.files_by_name[f.name].push_back(index);
			if (f.digest) out.files_by_digest[*f.digest].push_back(index);
			out.files.push_back(std::move(f));
		}
		// unknown kinds are skipped (decode/directory.rs:76-79)
	}
}

// ================================================================================================
// trailer.rs
static Bytes epilogue_bytes(const Trailer& t, uint8_t check) {
	Bytes b;
	b.push_back((uint8_t)t.digest_type);
	uint64_t off = (uint64_t)t.directory_offset;
	for (int i = 0; i < 8; i++) b.push_back((uint8_t)(off >> (8 * i)));
	for (int i = 0; i < 8; i++) b.push_back((uint8_t)(t.directory_uncompressed_size >> (8 * i)));
	b.push_back(check);
	b.push_back(t.version);
	b.insert(b.end(), ZARC_MAGIC, ZARC_MAGIC + 3);
	return b;
}
uint8_t Trailer::compute_check() const {  // XOR over prologue [0, digest_type], digest, epilogue with check = 0
	uint8_t c = 0 ^ (uint8_t)digest_type;
	for (uint8_t x : digest.bytes) c ^= x;
	for (uint8_t x : epilogue_bytes(*this, 0)) c ^= x;
	return c;
}
Bytes Trailer::to_bytes() const {
	Bytes b = digest.bytes;
	Bytes e = epilogue_bytes(*this, compute_check());
	b.insert(b.end(), e.begin(), e.end());
	return b;
}
void Trailer::make_offset_positive(uint64_t file_length) {
	if (directory_offset < 0) directory_offset += (int64_t)file_length;
}

// ================================================================================================
// staging buffer for the host-buffer ABI: page-locked when large (the copies then overlap the kernels),
// plain memory when small (locking pages costs more than it saves on a one-pass job)
struct HostBuffer {
	uint8_t* p = nullptr;
	bool pinned = false;
	std::unique_ptr<uint8_t[]> plain;
	explicit HostBuffer(uint64_t n) {
		if (n > (1ull << 30)) {
			p = (uint8_t*)gpu().zg_alloc_pinned(n + 1);
			pinned = p != nullptr;
		}
		if (!p) {
			plain.reset(new uint8_t[n + 1]);  // not zero-filled
			p = plain.get();
		}
	}
	~HostBuffer() {
		if (pinned) gpu().zg_free_pinned(p);
	}
	HostBuffer(const HostBuffer&) = delete;
};

// ================================================================================================
// Encoder
Encoder::Encoder(std::FILE* writer) : writer_(writer) {
	GpuLib& g = gpu();
	cctx_ = g.zg_cctx_create();
	if (!cctx_) throw Error("failed allocating zstd context");  // encode.rs:60-61 (also: no CUDA device)
	g.check(g.zg_cctx_init((zg_cctx*)cctx_, 0), "zstd init");
	write_all(FILE_MAGIC, sizeof FILE_MAGIC);
	offset_ = sizeof FILE_MAGIC;
	g.check(g.zg_cctx_reset_archive((zg_cctx*)cctx_, offset_), "archive reset");
}
Encoder::~Encoder() {
	if (cctx_) gpu().zg_cctx_free((zg_cctx*)cctx_);
}
void Encoder::write_all(const void* p, size_t n) {
	if (n && fwrite(p, 1, n, writer_) != n) throw Error(std::string("write failed: ") + strerror(errno));
}
void Encoder::set_zstd_parameter(ZstdParameter p, int value) {
	GpuLib& g = gpu();
	g.check(g.zg_cctx_set_parameter((zg_cctx*)cctx_, (int)p, value), "zstd parameter");
}
void Encoder::enable_compression(bool compress) {
	// the reference's store path writes frames libzstd itself rejects; refused here
	if (!compress) throw Error("--store is not supported: the reference's uncompressed frames are not valid Zstandard");
}

std::vector<Digest> Encoder::add_data_frames(const uint8_t* blob, const uint64_t* off, const uint64_t* len, size_t n) {
	std::vector<Digest> out(n);
	if (n == 0) return out;
	GpuLib& g = gpu();
	uint64_t cap = 0;
	for (size_t i = 0; i < n; i++) cap += len[i] + std::max<uint64_t>(1024, len[i] / 10);  // lowlevel_frames.rs:21
	std::vector<uint8_t> digests(n * DIGEST_LEN), first(n);
	std::vector<uint64_t> foff(n), flen(n);
	HostBuffer fb(cap);
	uint8_t* frames = fb.p;
	uint64_t bytes = 0;
	g.check(g.zg_pack_batch((zg_cctx*)cctx_, blob, off, len, n, digests.data(), first.data(), foff.data(), flen.data(), frames, cap, &bytes), "compress");
	write_all(frames, bytes);
	offset_ += bytes;
	for (size_t i = 0; i < n; i++) {
		out[i].bytes.assign(digests.begin() + i * DIGEST_LEN, digests.begin() + (i + 1) * DIGEST_LEN);
		if (first[i]) {  // content_frame.rs:48-57
			Frame f;
			f.edition = edition_;
			f.offset = foff[i];
			f.digest = out[i];
			f.length = flen[i];
			f.uncompressed = len[i];
			frames_.emplace(out[i], std::move(f));
			frame_order_.push_back(out[i]);
		}
	}
	return out;
}
Digest Encoder::add_data_frame(const uint8_t* content, size_t n) {
	static const uint8_t empty = 0;
	uint64_t off = 0, len = n;
	return add_data_frames(content ? content : &empty, &off, &len, 1)[0];
}

File Encoder::build_file(const Pathname& name) const {
	File f;
	f.edition = edition_;
	f.name = name;
	return f;
}

static std::optional<PosixOwner> owner_from(bool is_user, uint32_t id) {
	// posix_owner.rs from_uid/from_gid: id plus the name when the database knows it (cached per thread there)
	static thread_local std::map<std::pair<bool, uint32_t>, std::optional<std::string>> cache;
	auto key = std::make_pair(is_user, id);
	auto it = cache.find(key);
	if (it == cache.end()) {
		std::optional<std::string> name;
		char buf[4096];
		if (is_user) {
			passwd pw, *res = nullptr;
			if (getpwuid_r(id, &pw, buf, sizeof buf, &res) == 0 && res) name = res->pw_name;
		} else {
			group gr, *res = nullptr;
			if (getgrgid_r(id, &gr, buf, sizeof buf, &res) == 0 && res) name = res->gr_name;
		}
		it = cache.emplace(key, name).first;
	}
	PosixOwner o;
	o.id = id;
	if (it->second) o.name = CborString::from_maybe_utf8(*it->second);
	return o;
}

File Encoder::build_file_with_metadata(const std::string& path, bool follow_symlinks) const {
	File f = build_file(Pathname::from_normal_components(path));
	struct stat sym;
	if (lstat(path.c_str(), &sym) != 0) throw Error(path + ": " + strerror(errno));
	bool is_symlink = S_ISLNK(sym.st_mode);
	std::string target;
	if (is_symlink) {
		char buf[PATH_MAX];
		ssize_t k = readlink(path.c_str(), buf, sizeof buf);
		if (k < 0) throw Error(path + ": " + strerror(errno));
		target.assign(buf, (size_t)k);
	}
	struct stat st = sym;
	if (follow_symlinks && is_symlink && stat(path.c_str(), &st) != 0) throw Error(path + ": " + strerror(errno));
	f.user = owner_from(true, st.st_uid);
	f.group = owner_from(false, st.st_gid);
	f.mode = st.st_mode;
	if (S_ISDIR(st.st_mode)) {
		SpecialFile s;
		s.kind = SpecialFileKind::Directory;
		f.special = s;
	} else if (is_symlink) {
		SpecialFile s;
		s.kind = SpecialFileKind::Symlink;
		LinkTarget t;  // specials.rs:139-151: absolute or non-normal paths whole, else components
		bool normal = !target.empty() && target[0] != '/';
		if (normal) {
			size_t i = 0;
			while (i <= target.size()) {
				size_t j = target.find('/', i);
				if (j == std::string::npos) j = target.size();
				std::string c = target.substr(i, j - i);
				if (c == "." || c == "..") normal = false;
				i = j + 1;
			}
		}
		if (normal) {
			t.is_components = true;
			t.components = Pathname::from_normal_components(target).components;
		} else t.full_path = CborString::from_maybe_utf8(target);
		s.link_target = std::move(t);
		f.special = std::move(s);
	}
	Timestamps ts;
	ts.modified = Timestamp{(int64_t)st.st_mtim.tv_sec, (uint32_t)st.st_mtim.tv_nsec};
	ts.accessed = Timestamp{(int64_t)st.st_atim.tv_sec, (uint32_t)st.st_atim.tv_nsec};
#ifdef STATX_BTIME
	struct statx sx;
	int flags = (follow_symlinks ? 0 : AT_SYMLINK_NOFOLLOW);
	if (statx(AT_FDCWD, path.c_str(), flags, STATX_BTIME, &sx) == 0 && (sx.stx_mask & STATX_BTIME))
		ts.created = Timestamp{(int64_t)sx.stx_btime.tv_sec, sx.stx_btime.tv_nsec};
#endif
	f.timestamps = ts;
	// file attributes (chattr flags) and extended attributes are metadata bookkeeping outside the content
	// path: not gathered
	return f;
}

void Encoder::add_file_entry(File entry) {
	if (entry.digest && entry.digest->bytes.size() != DIGEST_LEN) throw Error("file entry digest has the wrong length");
	if (entry.digest && !frames_.count(*entry.digest)) throw Error("cannot add file entry referencing unknown data frame");
	size_t index = files_.size();
	files_by_name_[entry.name].push_back(index);
	if (entry.digest) files_by_digest_[*entry.digest].push_back(index);
	files_.push_back(std::move(entry));
}

size_t Encoder::write_compressed_frame(const uint8_t* data, size_t n) {
	GpuLib& g = gpu();
	size_t cap = n + std::max<size_t>(1024, n / 10);
	Bytes buf(cap);
	size_t bytes = g.check(g.zg_compress2((zg_cctx*)cctx_, buf.data(), cap, data, n), "compress");
	write_all(buf.data(), bytes);
	offset_ += bytes;
	return bytes;
}
size_t Encoder::write_skippable_frame(uint8_t nibble, const Bytes& payload) {
	uint8_t head[8] = {(uint8_t)(0x50 | (nibble & 15)), 0x2A, 0x4D, 0x18, (uint8_t)payload.size(), (uint8_t)(payload.size() >> 8),
	                   (uint8_t)(payload.size() >> 16), (uint8_t)(payload.size() >> 24)};
	write_all(head, 8);
	write_all(payload.data(), payload.size());
	offset_ += 8 + payload.size();
	return 8 + payload.size();
}

Digest Encoder::finalise() {
	if (finalised_) throw Error("encoder already finalised");
	finalised_ = true;
	GpuLib& g = gpu();
	Bytes directory;
	auto emit = [&](const Bytes& b) { directory.insert(directory.end(), b.begin(), b.end()); };
	Edition ed;
	ed.number = edition_;
	ed.written_at = Timestamp::now();
	ed.digest_type = DigestType::Blake3;
	emit(encode_element(ed));
	for (auto& kv : files_by_name_) {
		for (size_t index : kv.second) {
			if (index >= files_.size() || !files_[index]) continue;
			File file = std::move(*files_[index]);
			files_[index].reset();
			if (file.digest) {  // the frame element goes before the first file that links it
				auto it = frames_.find(*file.digest);
				if (it != frames_.end()) {
					emit(encode_element(it->second));
					frames_.erase(it);
				}
			}
			emit(encode_element(file));
		}
	}
	for (const Digest& d : frame_order_) {  // frames no file links
		auto it = frames_.find(d);
		if (it == frames_.end()) continue;
		emit(encode_element(it->second));
		frames_.erase(it);
	}
	// the directory stream is hashed and compressed by the same GPU kernels as file contents
	Digest digest;
	digest.bytes.resize(DIGEST_LEN);
	g.check(g.zg_blake3(directory.data(), directory.size(), digest.bytes.data()), "digest");
	size_t bytes = write_compressed_frame(directory.data(), directory.size());
	Trailer trailer;
	trailer.digest = digest;
	trailer.directory_uncompressed_size = directory.size();
	trailer.directory_offset = -(int64_t)(bytes + SKIPPABLE_FRAME_OVERHEAD + trailer.len());
	write_skippable_frame(0xF, trailer.to_bytes());
	if (fflush(writer_) != 0) throw Error(std::string("flush failed: ") + strerror(errno));
	return digest;
}

// ================================================================================================
// Decoder
Decoder::~Decoder() {
	if (dctx_) gpu().zg_dctx_free((zg_dctx*)dctx_);
}
Decoder::Decoder(Decoder&& o) noexcept
    : path_(std::move(o.path_)), file_length_(o.file_length_), trailer_(std::move(o.trailer_)), dir_(std::move(o.dir_)), dctx_(o.dctx_) {
	o.dctx_ = nullptr;
}
Bytes Decoder::read_at(uint64_t off, uint64_t n) const {
	Bytes b(n);
	int fd = ::open(path_.c_str(), O_RDONLY);
	if (fd < 0) throw Error(path_ + ": " + strerror(errno));
	uint64_t done = 0;
	while (done < n) {
		ssize_t k = pread(fd, b.data() + done, n - done, (off_t)(off + done));
		if (k < 0) {
			int e = errno;
			close(fd);
			throw Error(path_ + ": " + strerror(e));
		}
		if (k == 0) break;
		done += (uint64_t)k;
	}
	close(fd);
	b.resize(done);
	return b;
}

Decoder Decoder::open(const std::string& path) {
	Decoder d;
	d.path_ = path;
	struct stat st;
	if (stat(path.c_str(), &st) != 0) throw Error(path + ": " + strerror(errno));
	d.file_length_ = (uint64_t)st.st_size;
	// header (open.rs:33-67)
	Bytes head = d.read_at(0, 12);
	if (head.size() < 12 || memcmp(head.data(), FILE_MAGIC, 8) != 0 || memcmp(head.data() + 8, ZARC_MAGIC, 3) != 0)
		throw Error("parse error: not a zarc file (header magic)");
	if (head[11] != ZARC_VERSION) throw Error("unsupported zarc version " + std::to_string(head[11]));
	// trailer: the last <= 1 KiB, epilogue at the very end (open.rs:71-133)
	uint64_t ending_len = std::min<uint64_t>(d.file_length_, 1024);
	Bytes ending = d.read_at(d.file_length_ - ending_len, ending_len);
	if (ending.size() < EPILOGUE_LENGTH + 2 + DIGEST_LEN) throw Error("parse error: file too short for a zarc trailer");
	const uint8_t* e = ending.data() + ending.size() - EPILOGUE_LENGTH;
	Trailer t;
	if (e[0] != (uint8_t)DigestType::Blake3) throw Error("parse error: unknown digest type");
	t.digest_type = DigestType::Blake3;
	uint64_t off = 0, usz = 0;
	for (int i = 0; i < 8; i++) off |= (uint64_t)e[1 + i] << (8 * i);
	for (int i = 0; i < 8; i++) usz |= (uint64_t)e[9 + i] << (8 * i);
	t.directory_offset = (int64_t)off;
	t.directory_uncompressed_size = usz;
	uint8_t check = e[17];
	t.version = e[18];
	if (memcmp(e + 19, ZARC_MAGIC, 3) != 0) throw Error("parse error: trailer magic");
	t.digest.bytes.assign(e - DIGEST_LEN, e);
	uint8_t want = t.compute_check();
	if (want != check) {
		char msg[96];
		snprintf(msg, sizeof msg, "parse error: trailer check byte doesn't match (expected 0x%02X, got 0x%02X)", check, want);
		throw Error(msg);
	}
	t.make_offset_positive(d.file_length_);
	d.trailer_ = std::move(t);
	d.dctx_ = gpu().zg_dctx_create();
	if (!d.dctx_) throw Error("failed allocating zstd context");
	return d;
}

void Decoder::read_directory() {
	GpuLib& g = gpu();
	uint64_t off = (uint64_t)trailer_.directory_offset;
	if (trailer_.directory_offset < 0 || off >= file_length_) throw Error("parse error: directory offset outside the file");
	Bytes src = read_at(off, file_length_ - off);
	size_t fsz = g.check(g.zg_find_frame_compressed_size(src.data(), src.size()), "directory frame");
	Bytes data(trailer_.directory_uncompressed_size ? trailer_.directory_uncompressed_size : 1);
	size_t got = g.check(g.zg_decompress((zg_dctx*)dctx_, data.data(), trailer_.directory_uncompressed_size, src.data(), fsz), "directory frame");
	data.resize(got);
	Digest d;
	d.bytes.resize(DIGEST_LEN);
	g.check(g.zg_blake3(data.data(), data.size(), d.bytes.data()), "digest");
	Directory dir;
	parse_directory_stream(data.data(), data.size(), dir);
	dir_ = std::move(dir);
	if (d != trailer_.digest) throw Error("directory integrity check failed: digest");  // decode/directory.rs:114
}

const Frame* Decoder::frame(const Digest& d) const {
	auto it = dir_.frames.find(d);
	return it == dir_.frames.end() ? nullptr : &it->second;
}

std::vector<ContentFrame> Decoder::read_content_frames(const std::vector<Digest>& digests) {
	GpuLib& g = gpu();
	size_t n = digests.size();
	std::vector<ContentFrame> out(n);
	if (n == 0) return out;
	std::vector<uint64_t> off(n), len(n), ulen(n);
	std::vector<uint8_t> want(n * DIGEST_LEN), ok(n);
	std::vector<uint32_t> status(n);
	uint64_t cbytes = 0, ubytes = 0;
	for (size_t i = 0; i < n; i++) {
		const Frame* f = frame(digests[i]);
		if (!f) throw Error("frame not found");
		if (f->offset > file_length_ || f->length > file_length_ - f->offset) throw Error("parse error: frame outside the file");
		if (f->digest.bytes.size() != DIGEST_LEN) throw Error("parse error: frame digest has the wrong length");
		// sizes come from the (untrusted) directory: the running totals must not wrap
		if (f->uncompressed > UINT64_MAX - ubytes || f->length > UINT64_MAX - cbytes) throw Error("parse error: frame sizes overflow");
		off[i] = cbytes;
		len[i] = f->length;
		ulen[i] = f->uncompressed;
		cbytes += f->length;
		ubytes += f->uncompressed;
		memcpy(want.data() + i * DIGEST_LEN, f->digest.bytes.data(), DIGEST_LEN);
	}
	HostBuffer abuf(cbytes), pbuf(ubytes);
	uint8_t* arch = abuf.p;
	uint8_t* plain = pbuf.p;
	std::string err;
	int fd = ::open(path_.c_str(), O_RDONLY);
	if (fd < 0) throw Error(path_ + ": " + strerror(errno));
	for (size_t i = 0; i < n; i++) {
		const Frame* f = frame(digests[i]);
		uint64_t done = 0;
		while (done < f->length) {
			ssize_t k = pread(fd, arch + off[i] + done, f->length - done, (off_t)(f->offset + done));
			if (k <= 0) {
				close(fd);
				throw Error(path_ + ": short read");
			}
			done += (uint64_t)k;
		}
	}
	close(fd);
	size_t r = g.zg_unpack_batch((zg_dctx*)dctx_, arch, cbytes, n, off.data(), len.data(), ulen.data(), want.data(), plain, ubytes, nullptr, ok.data(),
	                             status.data());
	if (g.zg_is_error(r)) {
		auto code = g.zg_get_error_code(r);
		if (code == ZG_error_device || code == ZG_error_memory_allocation || code == ZG_error_no_device) throw Error(g.zg_error_name(r));
	}
	uint64_t pos = 0;
	for (size_t i = 0; i < n; i++) {
		if (status[i]) throw Error(std::string("zstd: ") + g.zg_error_name((size_t)0 - status[i]));  // decode/error.rs:35-38
		out[i].data.assign(plain + pos, plain + pos + ulen[i]);
		pos += ulen[i];
		out[i].verified = ok[i] != 0;
	}
	// the digest of the decoded bytes (FrameIterator::digest): equal to the directory's when verified
	for (size_t i = 0; i < n; i++) {
		if (*out[i].verified) out[i].digest = digests[i];
		else {
			out[i].digest.bytes.resize(DIGEST_LEN);
			g.check(g.zg_blake3(out[i].data.data(), out[i].data.size(), out[i].digest.bytes.data()), "digest");
		}
	}
	return out;
}

std::optional<ContentFrame> Decoder::read_content_frame(const Digest& d) {
	if (!frame(d)) return std::nullopt;  // Ok(None), decode/frame_iterator.rs:20-22
	return std::move(read_content_frames({d})[0]);
}

}  // namespace zarc
