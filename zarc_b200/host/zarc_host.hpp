// C++ host side of the zarc content path: the Encoder / Decoder API of crates/zarc (encode.rs,
// decode.rs) and the container format around the content frames (header, directory, trailer),
// above the C ABI of libzarcgpu.so.  The reference's host language (Rust) is not available in this
// build environment, so the host mirror is C++; names, argument meaning and error behaviour follow
// the reference so that its call sites (zarc-cli pack.rs:219-272, unpack.rs:38-138,
// list_files.rs:34-63) translate line by line.  Byte layout: SURVEY.md App. A.
//
// Every content byte (file contents AND the directory stream) is hashed, compressed and restored by
// the CUDA library; the host only serialises metadata (CBOR), lays frames out in the file and keeps
// the maps the reference keeps.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace zarc {

using Bytes = std::vector<uint8_t>;

constexpr uint8_t ZARC_VERSION = 1;                   // constants.rs
constexpr uint8_t ZARC_MAGIC[3] = {0x65, 0xAA, 0xDC};  // constants.rs
extern const uint8_t FILE_MAGIC[12];                  // header.rs:35-40
constexpr size_t SKIPPABLE_FRAME_OVERHEAD = 8;
constexpr size_t EPILOGUE_LENGTH = 22;                // trailer.rs:139
constexpr size_t DIGEST_LEN = 32;                     // integrity.rs:98-104

struct Error : std::runtime_error {
	using std::runtime_error::runtime_error;
};

// ---- integrity.rs ----
struct Digest {
	Bytes bytes;
	bool operator==(const Digest& o) const;  // constant-time, like integrity.rs:17-22
	bool operator!=(const Digest& o) const { return !(*this == o); }
	bool operator<(const Digest& o) const { return bytes < o.bytes; }
	std::string base64() const;
	static Digest from_base64(const std::string& s);
};
struct DigestHash {
	size_t operator()(const Digest& d) const;
};
enum class DigestType : uint8_t { Blake3 = 1 };

// ---- directory/strings.rs ----
struct CborString {  // text when valid UTF-8, else bytes (strings.rs:73-82)
	bool is_text = true;
	std::string data;
	static CborString from_maybe_utf8(const std::string& raw);
	bool operator<(const CborString& o) const;
	bool operator==(const CborString& o) const { return is_text == o.is_text && data == o.data; }
};
struct Pathname {
	std::vector<CborString> components;
	static Pathname from_normal_components(const std::string& path);  // strings.rs:22-35
	std::string to_path() const;                                      // strings.rs:38-62
	std::optional<std::string> to_safe_path() const;                  // extraction: no "..", no absolute or '/'-holding components
	bool operator<(const Pathname& o) const { return components < o.components; }
	bool operator==(const Pathname& o) const { return components == o.components; }
};
struct AttributeValue {
	bool is_bool = false;
	bool b = false;
	CborString s;
};
using AttributeMap = std::map<std::string, AttributeValue>;

// ---- directory/timestamps.rs ----
struct Timestamp {  // UTC
	int64_t secs = 0;
	uint32_t nanos = 0;
	static Timestamp now();
	std::string to_rfc3339() const;
	static Timestamp parse_rfc3339(const std::string& s);
	bool operator==(const Timestamp& o) const { return secs == o.secs && nanos == o.nanos; }
};
struct Timestamps {
	std::optional<Timestamp> created, modified, accessed;
};

// ---- directory/posix_owner.rs, specials.rs ----
struct PosixOwner {
	std::optional<uint64_t> id;
	std::optional<CborString> name;
};
enum class SpecialFileKind : uint8_t {
	Directory = 1, Symlink = 10, InternalSymlink = 11, ExternalAbsoluteSymlink = 12, ExternalRelativeSymlink = 13,
	Hardlink = 20, InternalHardlink = 21, ExternalHardlink = 22,
};
struct LinkTarget {
	bool is_components = false;
	CborString full_path;
	std::vector<CborString> components;
};
struct SpecialFile {
	std::optional<SpecialFileKind> kind;
	std::optional<LinkTarget> link_target;
	bool is_dir() const { return kind && *kind == SpecialFileKind::Directory; }
	bool is_symlink() const { return kind && (uint8_t)*kind >= 10 && (uint8_t)*kind <= 13; }
	bool is_hardlink() const { return kind && (uint8_t)*kind >= 20 && (uint8_t)*kind <= 22; }
};

// ---- directory/{edition,file,frame}.rs ----
struct Edition {
	uint16_t number = 1;
	Timestamp written_at;
	DigestType digest_type = DigestType::Blake3;
	std::optional<AttributeMap> user_metadata;
};
struct File {
	uint16_t edition = 1;
	Pathname name;
	std::optional<Digest> digest;
	std::optional<uint32_t> mode;
	std::optional<PosixOwner> user, group;
	std::optional<Timestamps> timestamps;
	std::optional<SpecialFile> special;
	std::optional<AttributeMap> user_metadata, attributes, extended_attributes;
	bool is_normal() const { return digest.has_value() && !special.has_value(); }  // file.rs:64-67
	bool is_dir() const { return special && special->is_dir(); }
	bool is_symlink() const { return special && special->is_symlink(); }
	bool is_hardlink() const { return special && special->is_hardlink(); }
};
struct Frame {
	uint16_t edition = 1;
	uint64_t offset = 0;
	Digest digest;
	uint64_t length = 0;        // whole Zstandard frame
	uint64_t uncompressed = 0;
};

// ---- directory/elements.rs: kind u8 | len u16 LE | 00 | CBOR[len] ----
enum class ElementKind : uint8_t { Edition = 1, File = 2, Frame = 3 };
Bytes encode_element(const Edition&);
Bytes encode_element(const File&);
Bytes encode_element(const Frame&);
struct Directory {
	std::map<uint16_t, Edition> editions;
	std::vector<File> files;
	std::unordered_map<Digest, Frame, DigestHash> frames;
	std::map<Pathname, std::vector<size_t>> files_by_name;
	std::unordered_map<Digest, std::vector<size_t>, DigestHash> files_by_digest;
};
// parses the concatenated element stream (decode/directory.rs:65-104); unknown kinds are skipped
void parse_directory_stream(const uint8_t* p, size_t n, Directory& out);

// ---- trailer.rs ----
struct Trailer {
	Digest digest;
	DigestType digest_type = DigestType::Blake3;
	int64_t directory_offset = 0;
	uint64_t directory_uncompressed_size = 0;
	uint8_t version = ZARC_VERSION;
	size_t len() const { return digest.bytes.size() + EPILOGUE_LENGTH; }  // trailer.rs:79-81
	uint8_t compute_check() const;                                        // trailer.rs:98-108
	Bytes to_bytes() const;                                               // trailer.rs:66-75 (no prologue)
	void make_offset_positive(uint64_t file_length);
};

// ---- encode.rs ----
enum class ZstdParameter : int {  // zstd_safe::CParameter as the CLI can name it (pack.rs:140-195)
	CompressionLevel = 100, WindowLog = 101, HashLog = 102, ChainLog = 103, SearchLog = 104, MinMatch = 105, TargetLength = 106,
	Strategy = 107, ContentSizeFlag = 200, ChecksumFlag = 201, DictIdFlag = 202,
};

struct zg_cctx_deleter;
class Encoder {
public:
	// Encoder::new (encode.rs:58-78): creates the compression context, writes FILE_MAGIC
	explicit Encoder(std::FILE* writer);
	~Encoder();
	Encoder(const Encoder&) = delete;
	void set_zstd_parameter(ZstdParameter p, int value);  // encode.rs:84-89
	void enable_compression(bool compress);               // encode.rs:95-97 (false is refused: App. F #1)

	// add_data_frame (encode/content_frame.rs:20-60): digest -> dedup -> compress -> append -> Frame record
	Digest add_data_frame(const uint8_t* content, size_t n);
	// the same for an ordered batch, in one GPU pass: results equal n successive add_data_frame calls
	std::vector<Digest> add_data_frames(const uint8_t* blob, const uint64_t* off, const uint64_t* len, size_t n);

	File build_file(const Pathname& name) const;  // add_file.rs:49-65
	// metadata/encode.rs:19-77 (mode, owners, timestamps, directory / symlink specials)
	File build_file_with_metadata(const std::string& path, bool follow_symlinks) const;
	void add_file_entry(File entry);  // add_file.rs:22-46
	Digest finalise();                // encode/directory.rs:40-122

	uint64_t offset() const { return offset_; }
	size_t frame_count() const { return frames_.size(); }

private:
	void write_all(const void* p, size_t n);
	size_t write_compressed_frame(const uint8_t* data, size_t n);              // lowlevel_frames.rs:19-39
	size_t write_skippable_frame(uint8_t nibble, const Bytes& payload);        // lowlevel_frames.rs:91-107
	std::FILE* writer_;
	void* cctx_;
	uint16_t edition_ = 1;
	std::vector<std::optional<File>> files_;
	std::unordered_map<Digest, Frame, DigestHash> frames_;
	std::vector<Digest> frame_order_;  // insertion order (the reference's HashMap order is arbitrary)
	std::map<Pathname, std::vector<size_t>> files_by_name_;
	std::unordered_map<Digest, std::vector<size_t>, DigestHash> files_by_digest_;
	uint64_t offset_ = 0;
	bool finalised_ = false;
};

// ---- decode.rs ----
struct ContentFrame {  // what draining a FrameIterator yields (decode/frame_iterator.rs)
	Bytes data;
	Digest digest;                // of the decoded bytes
	std::optional<bool> verified;  // FrameIterator::verify (frame_iterator.rs:86-88)
};
class Decoder {
public:
	static Decoder open(const std::string& path);  // decode/open.rs:121-158: header, trailer, check byte
	void read_directory();                         // decode/directory.rs:55-119: decompress, parse, verify digest
	uint64_t file_length() const { return file_length_; }
	const Trailer& trailer() const { return trailer_; }
	const std::vector<File>& files() const { return dir_.files; }
	const Directory& directory() const { return dir_; }
	const Frame* frame(const Digest& d) const;
	// read_content_frame (decode/frame_iterator.rs:14-27) drained: nullopt when the digest is unknown
	std::optional<ContentFrame> read_content_frame(const Digest& d);
	// the same for many frames in one GPU pass (out[i].verified is false for a digest mismatch, and
	// a frame that fails to decode throws like the iterator's Err item would)
	std::vector<ContentFrame> read_content_frames(const std::vector<Digest>& digests);
	~Decoder();
	Decoder(Decoder&&) noexcept;

private:
	Decoder() = default;
	Bytes read_at(uint64_t off, uint64_t n) const;
	std::string path_;
	uint64_t file_length_ = 0;
	Trailer trailer_;
	Directory dir_;
	void* dctx_ = nullptr;
};

}  // namespace zarc
