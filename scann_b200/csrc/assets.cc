// assets.cc -- serialized-asset I/O in the reference's on-disk format (host side, C++).
//
// Host-side mirror of
//   ScannInterface::LoadArtifacts / Serialize      scann_ops/cc/scann.cc:105-264,504-601
//   ScannNumpy::Serialize (scann_assets.pbtxt)     scann_ops/cc/scann_npy.cc:272-282
//   .npy v1.0 reader/writer                        utils/io_npy.h:39-116
//   proto (de)serialisation                        utils/io_oss_wrapper.cc:55-79
//
// The image has no protoc / libprotobuf for C++, so this file carries a small proto2 wire
// codec and a text-format parser driven by a schema table that lists the fields of
// proto/scann.proto, partitioning.proto, hash.proto, projection.proto, exact_reordering.proto,
// brute_force.proto, distance_measure.proto, input_output.proto (PureDynamicConfig),
// auto_tuning.proto, partitioning/partitioner.proto, trees/kmeans_tree/kmeans_tree.proto,
// proto/centers.proto and data_format/features.proto (field numbers copied from those files).
#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/scann_b200.h"

namespace sbassets {

}  // namespace sbassets
namespace sb { void set_last_error(const char* msg); }  // index.cu: backs scann_b200_last_error()
namespace sbassets {
static int afail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  sb::set_last_error(buf);
  return code;
}

// ------------------------------------------------------------------------------------------
// schema
// ------------------------------------------------------------------------------------------
enum FT { T_INT32, T_INT64, T_UINT32, T_UINT64, T_BOOL, T_ENUM, T_FLOAT, T_DOUBLE, T_STRING, T_MSG };
struct Field { std::string name; int num; FT type; std::string tname; bool repeated; bool packed; };
struct Msg { std::vector<Field> fields; };
struct Schema {
  std::map<std::string, Msg> msgs;
  std::map<std::string, std::vector<std::pair<std::string, int>>> enums;
};

static const char kSchemaText[] = R"(
message ScannConfig { string dataset_name=32; int32 num_neighbors=3; float epsilon_distance=4;
  DistanceMeasureConfig distance_measure=5; ExactReordering exact_reordering=17; InputOutputConfig input_output=6;
  BruteForceConfig brute_force=7; PartitioningConfig partitioning=8; HashConfig hash=13;
  int32 num_single_shard_neighbors=21; AutopilotConfig autopilot=43; }
message DistanceMeasureConfig { string distance_measure=1; }
message ExactReordering { int32 approx_num_neighbors=1; float approx_epsilon_distance=2;
  DistanceMeasureConfig approx_distance_measure=3; FixedPoint fixed_point=5; Bfloat16 bfloat16=7;
  bool use_fixed_point_if_possible=4; }
message FixedPoint { bool enabled=1; float fixed_point_multiplier=2; string multipliers_filename=7;
  double noise_shaping_threshold=8; float fixed_point_multiplier_quantile=6; }
message Bfloat16 { bool enabled=1; double noise_shaping_threshold=2; }
message BruteForceConfig { bool scalar_quantized=1; FixedPoint fixed_point=4; Bfloat16 bfloat16=5;
  float scalar_quantization_multiplier_quantile=2; float scalar_quantization_noise_shaping_threshold=3; }
message InputOutputConfig { enum:InMemoryTypes in_memory_data_type=2; PureDynamicConfig pure_dynamic_config=21; }
message PureDynamicConfig { int32 num_shards=1; enum:VectorType vector_type=2; uint64 dimensionality=3; }
enum InMemoryTypes { INT8=0; UINT8=1; INT16=2; INT32=4; UINT32=5; INT64=6; FLOAT=8; DOUBLE=9; IN_MEMORY_DATA_TYPE_NOT_SPECIFIED=255; }
enum VectorType { UNSPECIFIED_VECTOR_TYPE=0; SPARSE=1; DENSE=2; }
message AutopilotConfig { AutopilotTreeAH tree_ah=1; }
message AutopilotTreeAH { int64 l1_size=1; int64 l3_size=2; enum:IncrementalMode incremental_mode=3;
  enum:AutopilotDataType reordering_dtype=4; int64 partitioning_expected_sample_size=5; int64 num_leaf_partitions=6;
  float fraction_leaf_partitions=8; int64 first_n_leaf_partitions_for_assignment=7; }
enum IncrementalMode { NONE=0; ONLINE=1; ONLINE_INCREMENTAL=2; }
enum AutopilotDataType { UNKNOWN=0; FLOAT32=1; BFLOAT16=2; INT8=3; }
message PartitioningConfig { enum:TreeType tree_type=31; float partitioning_sampling_fraction=4; int32 expected_sample_size=45;
  string partitioner_prefix=8; int32 clustering_seed=27; ProjectionConfig projection=12; int32 num_children=3;
  DistanceMeasureConfig partitioning_distance=10; DistanceMeasureConfig database_tokenization_distance_override=24;
  DistanceMeasureConfig query_tokenization_distance_override=25; enum:TokenizationType query_tokenization_type=28;
  enum:TokenizationType database_tokenization_type=29; int32 max_clustering_iterations=6; int32 num_mini_batches=38;
  bool ignore_empty_cluster_errors=55; float clustering_convergence_tolerance=7; float min_cluster_size=9;
  int32 max_cluster_size=40; double perturbation=41; enum:PartitioningType partitioning_type=23;
  enum:BalancingType balancing_type=35; enum:CenterInit single_machine_center_initialization=49;
  DatabaseSpillingConfig database_spilling=20; QuerySpillingConfig query_spilling=21; float avq=51;
  IncrementalTrainingConfig incremental_training_config=52; int32 num_tokenized_branch=53;
  BottomUpTopLevelPartitioner bottom_up_top_level_partitioner=54; int32 desired_average_cluster_size=34; }
enum TreeType { KMEANS_TREE=0; PCA_TREE=1; RANDOM_PROJECTION_TREE=2; BALL_TREE=3; RANDOM=4; TREE_X_HYBRID=5; }
enum TokenizationType { FLOAT=1; FIXED_POINT_INT8=2; ASYMMETRIC=3; }
enum PartitioningType { GENERIC=0; SPHERICAL=1; }
enum BalancingType { DEFAULT_UNBALANCED=0; GREEDY_BALANCED=1; UNBALANCED_FLOAT32=2; }
enum CenterInit { DEFAULT_KMEANS_PLUS_PLUS=0; RANDOM_INITIALIZATION=1; }
message DatabaseSpillingConfig { enum:DbSpillingType spilling_type=1; float replication_factor=2; uint32 max_spill_centers=3;
  float orthogonality_amplification_lambda=4; float overretrieve_factor=5; }
enum DbSpillingType { NO_SPILLING=0; MULTIPLICATIVE=1; ADDITIVE=2; FIXED_NUMBER_OF_CENTERS=3; TWO_CENTER_ORTHOGONALITY_AMPLIFIED=4; SOAR=4; }
message QuerySpillingConfig { enum:QuerySpillingType spilling_type=1; float spilling_threshold=2; uint32 max_spill_centers=3; }
enum QuerySpillingType { NO_SPILLING=0; MULTIPLICATIVE=1; ADDITIVE=2; ABSOLUTE_DISTANCE=3; FIXED_NUMBER_OF_CENTERS=4; }
message IncrementalTrainingConfig { float fraction=1; uint32 number_of_datapoints=2; uint32 cluster_stability_size=3;
  bool autopilot=4; uint32 max_split=5; }
message BottomUpTopLevelPartitioner { bool enabled=1; int32 num_centroids=2; int32 num_centroids_to_search=3; float avq=4;
  SoarConfig soar=5; enum:BottomUpQuantization quantization=6; float noise_shaping_threshold=7;
  BottomUpTopLevelPartitioner next_higher_level=8; }
message SoarConfig { bool enabled=1; float lambda=2; float overretrieve_factor=3; }
enum BottomUpQuantization { FLOAT32=0; BFLOAT16=1; FIXED8=2; }
message HashConfig { int32 num_bits=1; ProjectionConfig projection=2; string parameters_filename=4;
  AsymmetricHasherConfig asymmetric_hash=5; }
message AsymmetricHasherConfig { ProjectionConfig projection=1; int32 num_clusters_per_block=2; enum:LookupType lookup_type=20;
  int32 clustering_seed=9; bool use_residual_quantization=22; double noise_shaping_threshold=28;
  bool use_norm_biasing_correction=31; bool use_global_topn=33;
  FixedPointLUTConversionOptions fixed_point_lut_conversion_options=25; float sampling_fraction=10;
  int32 expected_sample_size=29; int32 sampling_seed=11; DistanceMeasureConfig quantization_distance=18;
  int32 max_clustering_iterations=4; float clustering_convergence_tolerance=5; string centers_filename=6;
  bool use_per_leaf_partition_training=17; enum:QuantizationScheme quantization_scheme=23; int32 max_sample_size=3;
  float min_cluster_size=19; }
enum LookupType { FLOAT=0; INT8=1; INT16=2; INT8_LUT16=3; }
enum QuantizationScheme { PRODUCT=0; STACKED=1; PRODUCT_AND_BIAS=2; PRODUCT_AND_PACK=3; }
message FixedPointLUTConversionOptions { enum:FloatToInt float_to_int_conversion_method=1; float multiplier_quantile=2; }
enum FloatToInt { TRUNCATE=0; ROUND=1; }
message ProjectionConfig { enum:ProjectionType projection_type=1; uint64 input_dim=9; int32 num_blocks=2;
  int32 num_dims_per_block=3; repeated VariableBlock variable_blocks=4; int32 seed=5; bool is_dense=7;
  bool build_covariance=8; float pca_significance_threshold=13; float pca_truncation_threshold=14;
  bool pca_random_rotate_projection_matrix=15; }
message VariableBlock { int32 num_blocks=1; int32 num_dims_per_block=2; }
enum ProjectionType { NONE=0; CHUNK=1; VARIABLE_CHUNK=2; RANDOM_GAUSS=3; RANDOM_BINARY=4; RANDOM_BINARY_DYNAMIC=5;
  RANDOM_SPARSE_BINARY=6; RANDOM_ORTHOGONAL=7; PCA=8; RANDOM_BILINEAR=9; MEANSTD_PROJECTION=12; IDENTITY_CHUNK=13;
  TRUNCATE=14; EIGENVALUE_OPQ=15; }
message ScannAssets { repeated ScannAsset assets=1; bool trained_on_the_fly=2; }
message ScannAsset { enum:AssetType asset_type=1; string asset_path=2; }
enum AssetType { UNSPECIFIED_TYPE=0; DATASET=1; INT8_DATASET=2; AH_DATASET=3; TOKENIZATION=4;
  REORDERING_INT8_MULTIPLIERS=5; BRUTE_FORCE_INT8_MULTIPLIERS=6; AH_CENTERS=7; PARTITIONER=8; DATASET_NPY=9;
  INT8_DATASET_NPY=10; AH_DATASET_NPY=11; AH_DATASET_SOAR_NPY=16; TOKENIZATION_NPY=12; INT8_MULTIPLIERS_NPY=13;
  INT8_NORMS_NPY=14; BF16_DATASET_NPY=15; }
)";

static const Schema& schema() {
  static Schema* s = [] {
    Schema* sc = new Schema();
    std::istringstream in(kSchemaText);
    std::string tok;
    auto next = [&](std::string& t) -> bool {
      t.clear();
      int c;
      while ((c = in.get()) != EOF && (isspace(c))) {}
      if (c == EOF) return false;
      if (c == '{' || c == '}' || c == ';' || c == '=') { t = (char)c; return true; }
      t += (char)c;
      while ((c = in.peek()) != EOF && !isspace(c) && c != '{' && c != '}' && c != ';' && c != '=') t += (char)in.get();
      return true;
    };
    while (next(tok)) {
      std::string name;
      if (tok == "message") {
        next(name); next(tok);  // {
        Msg m;
        while (next(tok) && tok != "}") {
          Field f{};
          if (tok == "repeated") { f.repeated = true; next(tok); }
          std::string ty = tok;
          next(f.name); next(tok);  // =
          next(tok); f.num = atoi(tok.c_str());
          next(tok);  // ;
          if (ty == "int32") f.type = T_INT32; else if (ty == "int64") f.type = T_INT64;
          else if (ty == "uint32") f.type = T_UINT32; else if (ty == "uint64") f.type = T_UINT64;
          else if (ty == "bool") f.type = T_BOOL; else if (ty == "float") f.type = T_FLOAT;
          else if (ty == "double") f.type = T_DOUBLE; else if (ty == "string") f.type = T_STRING;
          else if (ty.rfind("enum:", 0) == 0) { f.type = T_ENUM; f.tname = ty.substr(5); }
          else { f.type = T_MSG; f.tname = ty; }
          m.fields.push_back(f);
        }
        sc->msgs[name] = m;
      } else if (tok == "enum") {
        next(name); next(tok);
        std::vector<std::pair<std::string, int>> vals;
        while (next(tok) && tok != "}") {
          std::string vn = tok;
          next(tok); next(tok);
          vals.push_back({vn, atoi(tok.c_str())});
          next(tok);
        }
        sc->enums[name] = vals;
      }
    }
    return sc;
  }();
  return *s;
}

// ------------------------------------------------------------------------------------------
// generic message tree + text format (protobuf TextFormat subset: `f: v`, `f { }`, `f: { }`,
// `f < >`, comments, Python-style True/False/nan/inf literals -- scann_builder.py:213-238)
// ------------------------------------------------------------------------------------------
struct Node {
  std::string name;
  bool is_msg = false;
  std::string scalar;
  bool quoted = false;
  std::vector<Node> kids;
  const Node* find(const std::string& n) const {
    for (const auto& k : kids) if (k.name == n) return &k;
    return nullptr;
  }
  const Node* path(std::initializer_list<const char*> p) const {
    const Node* cur = this;
    for (const char* n : p) { if (!cur) return nullptr; cur = cur->find(n); }
    return cur;
  }
};

struct TextParser {
  const std::string& s;
  size_t i = 0;
  std::string err;
  explicit TextParser(const std::string& t) : s(t) {}
  void skip() {
    for (;;) {
      while (i < s.size() && (isspace((unsigned char)s[i]) || s[i] == ',' || s[i] == ';')) ++i;
      if (i < s.size() && s[i] == '#') { while (i < s.size() && s[i] != '\n') ++i; continue; }
      break;
    }
  }
  bool parse_fields(Node* out, char closer) {
    for (;;) {
      skip();
      if (i >= s.size()) { if (closer) { err = "unexpected end of text proto"; return false; } return true; }
      if (closer && s[i] == closer) { ++i; return true; }
      size_t b = i;
      while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '.')) ++i;
      if (b == i) { err = std::string("unexpected character '") + s[i] + "' in text proto"; return false; }
      Node n;
      n.name = s.substr(b, i - b);
      skip();
      bool colon = false;
      if (i < s.size() && s[i] == ':') { colon = true; ++i; skip(); }
      if (i < s.size() && (s[i] == '{' || s[i] == '<')) {
        const char close = s[i] == '{' ? '}' : '>';
        ++i;
        n.is_msg = true;
        if (!parse_fields(&n, close)) return false;
      } else {
        if (!colon) { err = "expected ':' or '{' after field " + n.name; return false; }
        if (i < s.size() && (s[i] == '"' || s[i] == '\'')) {
          const char q = s[i++];
          n.quoted = true;
          while (i < s.size() && s[i] != q) {
            if (s[i] == '\\' && i + 1 < s.size()) {
              ++i;
              char c = s[i];
              n.scalar += c == 'n' ? '\n' : c == 't' ? '\t' : c;
            } else n.scalar += s[i];
            ++i;
          }
          if (i >= s.size()) { err = "unterminated string"; return false; }
          ++i;
        } else {
          size_t vb = i;
          while (i < s.size() && !isspace((unsigned char)s[i]) && s[i] != ',' && s[i] != ';' && s[i] != '}' && s[i] != '>' && s[i] != '#') ++i;
          n.scalar = s.substr(vb, i - vb);
          if (n.scalar.empty()) { err = "missing value for field " + n.name; return false; }
        }
      }
      out->kids.push_back(std::move(n));
    }
  }
};

static bool parse_text(const std::string& text, Node* root, std::string* err) {
  TextParser p(text);
  root->is_msg = true;
  if (!p.parse_fields(root, 0)) { *err = p.err; return false; }
  return true;
}

// ---- wire helpers -------------------------------------------------------------------------
static void put_varint(std::string* o, uint64_t v) {
  while (v >= 0x80) { o->push_back((char)(v | 0x80)); v >>= 7; }
  o->push_back((char)v);
}
static void put_tag(std::string* o, int num, int wt) { put_varint(o, ((uint64_t)num << 3) | wt); }
static void put_f32(std::string* o, float f) { o->append(reinterpret_cast<const char*>(&f), 4); }
static void put_f64(std::string* o, double f) { o->append(reinterpret_cast<const char*>(&f), 8); }
static void put_len(std::string* o, int num, const std::string& payload) {
  put_tag(o, num, 2);
  put_varint(o, payload.size());
  o->append(payload);
}
struct Reader {
  const uint8_t* p; const uint8_t* e; bool ok = true;
  Reader(const void* b, size_t n) : p((const uint8_t*)b), e((const uint8_t*)b + n) {}
  bool done() const { return p >= e || !ok; }
  uint64_t varint() {
    uint64_t v = 0; int sh = 0;
    while (p < e) { uint8_t b = *p++; v |= (uint64_t)(b & 0x7F) << sh; if (!(b & 0x80)) return v; sh += 7; if (sh > 63) break; }
    ok = false; return 0;
  }
  float f32() { float f = 0; if (e - p < 4) { ok = false; return 0; } memcpy(&f, p, 4); p += 4; return f; }
  double f64() { double f = 0; if (e - p < 8) { ok = false; return 0; } memcpy(&f, p, 8); p += 8; return f; }
  Reader sub() { uint64_t n = varint(); if (!ok || (uint64_t)(e - p) < n) { ok = false; return Reader(p, 0); } Reader r(p, n); p += n; return r; }
  void skip(int wt) {
    if (wt == 0) varint(); else if (wt == 1) { if (e - p < 8) ok = false; else p += 8; }
    else if (wt == 2) sub(); else if (wt == 5) { if (e - p < 4) ok = false; else p += 4; } else ok = false;
  }
};

static bool parse_bool(const std::string& v, bool* out) {
  if (v == "true" || v == "True" || v == "t" || v == "1") { *out = true; return true; }
  if (v == "false" || v == "False" || v == "f" || v == "0") { *out = false; return true; }
  return false;
}
static bool parse_double(const std::string& v, double* out) {
  std::string t = v;
  for (auto& c : t) c = (char)tolower(c);
  if (!t.empty() && t.back() == 'f' && t != "inf" && t != "-inf" && t != "+inf") t.pop_back();
  if (t == "nan") { *out = NAN; return true; }
  if (t == "inf" || t == "infinity" || t == "+inf") { *out = INFINITY; return true; }
  if (t == "-inf" || t == "-infinity") { *out = -INFINITY; return true; }
  char* end = nullptr;
  *out = strtod(t.c_str(), &end);
  return end && *end == 0 && !t.empty();
}

static bool encode_msg(const Node& n, const std::string& mname, std::string* out, std::string* err) {
  const Schema& sc = schema();
  auto it = sc.msgs.find(mname);
  if (it == sc.msgs.end()) { *err = "unknown message type " + mname; return false; }
  for (const Node& k : n.kids) {
    const Field* f = nullptr;
    for (const Field& c : it->second.fields) if (c.name == k.name) { f = &c; break; }
    if (!f) { *err = "unknown field '" + k.name + "' in " + mname; return false; }
    if ((f->type == T_MSG) != k.is_msg) { *err = "field '" + k.name + "' of " + mname + ": message/scalar mismatch"; return false; }
    switch (f->type) {
      case T_MSG: {
        std::string sub;
        if (!encode_msg(k, f->tname, &sub, err)) return false;
        put_len(out, f->num, sub);
        break;
      }
      case T_STRING: put_len(out, f->num, k.scalar); break;
      case T_BOOL: {
        bool b;
        if (!parse_bool(k.scalar, &b)) { *err = "bad bool '" + k.scalar + "' for " + k.name; return false; }
        put_tag(out, f->num, 0); put_varint(out, b ? 1 : 0);
        break;
      }
      case T_ENUM: {
        const auto& vals = sc.enums.at(f->tname);
        int v = -1; bool found = false;
        for (const auto& e : vals) if (e.first == k.scalar) { v = e.second; found = true; break; }
        if (!found) {
          char* end = nullptr; long lv = strtol(k.scalar.c_str(), &end, 10);
          if (end && *end == 0 && !k.scalar.empty()) { v = (int)lv; found = true; }
        }
        if (!found) { *err = "bad enum value '" + k.scalar + "' for " + k.name; return false; }
        put_tag(out, f->num, 0); put_varint(out, (uint64_t)(int64_t)v);
        break;
      }
      case T_INT32: case T_INT64: case T_UINT32: case T_UINT64: {
        char* end = nullptr;
        errno = 0;
        long long v = strtoll(k.scalar.c_str(), &end, 0);
        unsigned long long uv = (unsigned long long)v;
        if (f->type == T_UINT64 || f->type == T_UINT32) uv = strtoull(k.scalar.c_str(), &end, 0);
        if (!end || *end != 0 || k.scalar.empty()) { *err = "bad integer '" + k.scalar + "' for " + k.name; return false; }
        put_tag(out, f->num, 0); put_varint(out, (uint64_t)uv);
        break;
      }
      case T_FLOAT: case T_DOUBLE: {
        double d;
        if (!parse_double(k.scalar, &d)) { *err = "bad number '" + k.scalar + "' for " + k.name; return false; }
        if (f->type == T_FLOAT) { put_tag(out, f->num, 5); put_f32(out, (float)d); }
        else { put_tag(out, f->num, 1); put_f64(out, d); }
        break;
      }
    }
  }
  return true;
}

static void fmt_float(std::ostringstream& o, double d, bool is_float) {
  if (isnan(d)) o << "nan"; else if (isinf(d)) o << (d > 0 ? "inf" : "-inf");
  else {  // shortest representation that round-trips, like protobuf's TextFormat printer
    char b[64];
    if (is_float) {
      for (int p = 6; p <= 9; ++p) { snprintf(b, sizeof b, "%.*g", p, d); if ((float)strtod(b, nullptr) == (float)d) break; }
    } else {
      for (int p = 15; p <= 17; ++p) { snprintf(b, sizeof b, "%.*g", p, d); if (strtod(b, nullptr) == d) break; }
    }
    o << b;
  }
}

static bool decode_msg(Reader r, const std::string& mname, int indent, std::ostringstream& o) {
  const Schema& sc = schema();
  auto it = sc.msgs.find(mname);
  if (it == sc.msgs.end()) return false;
  const std::string pad((size_t)indent * 2, ' ');
  while (!r.done()) {
    const uint64_t tag = r.varint();
    if (!r.ok) return false;
    const int num = (int)(tag >> 3), wt = (int)(tag & 7);
    const Field* f = nullptr;
    for (const Field& c : it->second.fields) if (c.num == num) { f = &c; break; }
    if (!f) { r.skip(wt); if (!r.ok) return false; continue; }
    switch (f->type) {
      case T_MSG: {
        if (wt != 2) return false;
        Reader s = r.sub();
        if (!r.ok) return false;
        o << pad << f->name << " {\n";
        if (!decode_msg(s, f->tname, indent + 1, o)) return false;
        o << pad << "}\n";
        break;
      }
      case T_STRING: {
        if (wt != 2) return false;
        Reader s = r.sub();
        if (!r.ok) return false;
        o << pad << f->name << ": \"";
        for (const uint8_t* c = s.p; c < s.e; ++c) { if (*c == '"' || *c == '\\') o << '\\'; o << (char)*c; }
        o << "\"\n";
        break;
      }
      case T_BOOL: o << pad << f->name << ": " << (r.varint() ? "true" : "false") << "\n"; break;
      case T_ENUM: {
        const int v = (int)(int64_t)r.varint();
        const auto& vals = sc.enums.at(f->tname);
        const char* nm = nullptr;
        for (const auto& e : vals) if (e.second == v) { nm = e.first.c_str(); break; }
        o << pad << f->name << ": ";
        if (nm) o << nm; else o << v;
        o << "\n";
        break;
      }
      case T_INT32: o << pad << f->name << ": " << (int32_t)(int64_t)r.varint() << "\n"; break;
      case T_INT64: o << pad << f->name << ": " << (int64_t)r.varint() << "\n"; break;
      case T_UINT32: o << pad << f->name << ": " << (uint32_t)r.varint() << "\n"; break;
      case T_UINT64: o << pad << f->name << ": " << (uint64_t)r.varint() << "\n"; break;
      case T_FLOAT: { if (wt != 5) return false; o << pad << f->name << ": "; fmt_float(o, r.f32(), true); o << "\n"; break; }
      case T_DOUBLE: { if (wt != 1) return false; o << pad << f->name << ": "; fmt_float(o, r.f64(), false); o << "\n"; break; }
    }
    if (!r.ok) return false;
  }
  return r.ok;
}

int text_to_binary(const char* mname, const std::string& text, std::string* out) {
  Node root;
  std::string err;
  if (!parse_text(text, &root, &err)) return afail(SCANN_B200_INVALID_ARGUMENT, "Failed to parse text proto: %s", err.c_str());
  out->clear();
  if (!encode_msg(root, mname, out, &err)) return afail(SCANN_B200_INVALID_ARGUMENT, "Failed to parse %s: %s", mname, err.c_str());
  return 0;
}
int binary_to_text(const char* mname, const std::string& bin, std::string* out) {
  std::ostringstream o;
  if (!decode_msg(Reader(bin.data(), bin.size()), mname, 0, o)) return afail(SCANN_B200_INVALID_ARGUMENT, "malformed %s protobuf", mname);
  *out = o.str();
  return 0;
}

// ------------------------------------------------------------------------------------------
// files
// ------------------------------------------------------------------------------------------
static int read_file(const std::string& path, std::string* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return afail(SCANN_B200_FAILED_PRECONDITION, "Failed to open file %s: %s", path.c_str(), strerror(errno));
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out->resize((size_t)n);
  size_t got = n ? fread(&(*out)[0], 1, (size_t)n, f) : 0;
  fclose(f);
  if (got != (size_t)n) return afail(SCANN_B200_INTERNAL, "short read on %s", path.c_str());
  return 0;
}
static int write_file(const std::string& path, const void* data, size_t n) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return afail(SCANN_B200_FAILED_PRECONDITION, "Failed to open file %s for writing: %s", path.c_str(), strerror(errno));
  size_t put = n ? fwrite(data, 1, n, f) : 0;
  int rc = fclose(f);
  if (put != n || rc != 0) return afail(SCANN_B200_INTERNAL, "short write on %s", path.c_str());
  return 0;
}

// utils/io_npy.h:39-73: magic, v1.0, u16 header length, dict padded so data starts at a multiple of 64.
static int write_npy(const std::string& path, const char* descr, const void* data, size_t elem, size_t n0, long n1) {
  char dict[256];
  if (n1 >= 0) snprintf(dict, sizeof dict, "{'descr': '%s', 'fortran_order': False, 'shape': (%zu, %ld), }", descr, n0, n1);
  else snprintf(dict, sizeof dict, "{'descr': '%s', 'fortran_order': False, 'shape': (%zu,), }", descr, n0);
  std::string h(dict);
  const size_t pre = 10;
  size_t total = pre + h.size() + 1;
  const size_t padded = (total + 63) / 64 * 64;
  h.append(padded - total, ' ');
  h.push_back('\n');
  std::string head("\x93NUMPY\x01\x00", 8);
  const uint16_t hl = (uint16_t)h.size();
  head.push_back((char)(hl & 0xFF));
  head.push_back((char)(hl >> 8));
  head += h;
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return afail(SCANN_B200_FAILED_PRECONDITION, "Failed to open file %s for writing: %s", path.c_str(), strerror(errno));
  const size_t cnt = n0 * (size_t)(n1 >= 0 ? n1 : 1);
  bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
  ok = ok && (cnt == 0 || fwrite(data, elem, cnt, f) == cnt);
  ok = (fclose(f) == 0) && ok;
  return ok ? 0 : afail(SCANN_B200_INTERNAL, "short write on %s", path.c_str());
}

struct NpyArray { std::string bytes; std::string descr; std::vector<size_t> shape; size_t data_off = 0; size_t elem = 0; };
static int read_npy(const std::string& path, NpyArray* a) {
  if (int rc = read_file(path, &a->bytes)) return rc;
  const std::string& b = a->bytes;
  if (b.size() < 10 || memcmp(b.data(), "\x93NUMPY", 6) != 0) return afail(SCANN_B200_INVALID_ARGUMENT, "%s is not an .npy file", path.c_str());
  const int major = (uint8_t)b[6];
  size_t hl, off;
  if (major == 1) { hl = (uint8_t)b[8] | ((size_t)(uint8_t)b[9] << 8); off = 10; }
  else { if (b.size() < 12) return afail(SCANN_B200_INVALID_ARGUMENT, "truncated npy"); hl = (uint8_t)b[8] | ((size_t)(uint8_t)b[9] << 8) | ((size_t)(uint8_t)b[10] << 16) | ((size_t)(uint8_t)b[11] << 24); off = 12; }
  if (off + hl > b.size()) return afail(SCANN_B200_INVALID_ARGUMENT, "truncated npy header in %s", path.c_str());
  const std::string h = b.substr(off, hl);
  size_t p = h.find("'descr'");
  if (p == std::string::npos) return afail(SCANN_B200_INVALID_ARGUMENT, "npy header without descr");
  p = h.find('\'', h.find(':', p));
  size_t q = h.find('\'', p + 1);
  a->descr = h.substr(p + 1, q - p - 1);
  if (h.find("'fortran_order': True") != std::string::npos) return afail(SCANN_B200_INVALID_ARGUMENT, "Fortran-ordered npy is not supported (%s)", path.c_str());
  p = h.find('(', h.find("'shape'"));
  q = h.find(')', p);
  std::string sh = h.substr(p + 1, q - p - 1);
  a->shape.clear();
  const char* c = sh.c_str();
  while (*c) {
    while (*c && !isdigit((unsigned char)*c)) ++c;
    if (!*c) break;
    a->shape.push_back((size_t)strtoull(c, const_cast<char**>(&c), 10));
  }
  a->elem = (size_t)atoi(a->descr.c_str() + 2);
  a->data_off = off + hl;
  size_t cnt = 1;
  for (size_t s : a->shape) cnt *= s;
  if (a->data_off + cnt * a->elem > b.size()) return afail(SCANN_B200_INVALID_ARGUMENT, "npy data truncated in %s", path.c_str());
  return 0;
}

// ------------------------------------------------------------------------------------------
// assets
// ------------------------------------------------------------------------------------------
struct Assets {
  std::string config_text;
  Node config;
  std::vector<float> centers, codebook, dataset_own;
  std::vector<int32_t> block_dims;
  NpyArray tokens, codes, soar_codes, dataset, bf16, int8, int8_mult, dp_norms;
  uint32_t n_leaves = 0, n_blocks = 0, dpb = 0, d = 0, n = 0;
  bool has_tokens = false, has_codes = false, has_soar = false, has_dataset = false, has_bf16 = false;
  bool has_int8 = false, has_int8_mult = false, has_dp_norms = false;
};

static double node_num(const Node* n, double dflt) {
  double d;
  if (n && !n->is_msg && parse_double(n->scalar, &d)) return d;
  return dflt;
}

// trees/kmeans_tree/kmeans_tree_node.cc:91-124: flat tree = root with L centers (double `dimension`
// or float `float_dimension`) and L leaf children.
static int parse_partitioner(const std::string& bin, Assets* a) {
  Reader r(bin.data(), bin.size());
  int n_tokens = -1;
  std::vector<std::vector<float>> centers;
  while (!r.done()) {
    uint64_t tag = r.varint(); int num = (int)(tag >> 3), wt = (int)(tag & 7);
    if (num == 1 && wt == 0) n_tokens = (int)r.varint();
    else if (num == 2 && wt == 2) {  // SerializedKMeansTreePartitioner
      Reader km = r.sub();
      while (!km.done()) {
        uint64_t t2 = km.varint(); int n2 = (int)(t2 >> 3), w2 = (int)(t2 & 7);
        if (n2 == 1 && w2 == 2) {  // SerializedKMeansTree
          Reader tree = km.sub();
          while (!tree.done()) {
            uint64_t t3 = tree.varint(); int n3 = (int)(t3 >> 3), w3 = (int)(t3 & 7);
            if (n3 == 1 && w3 == 2) {  // root Node
              Reader node = tree.sub();
              while (!node.done()) {
                uint64_t t4 = node.varint(); int n4 = (int)(t4 >> 3), w4 = (int)(t4 & 7);
                if (n4 == 1 && w4 == 2) {  // Center
                  Reader c = node.sub();
                  std::vector<float> v;
                  while (!c.done()) {
                    uint64_t t5 = c.varint(); int n5 = (int)(t5 >> 3), w5 = (int)(t5 & 7);
                    if (n5 == 1 && w5 == 2) { Reader p = c.sub(); while (!p.done()) v.push_back((float)p.f64()); }
                    else if (n5 == 1 && w5 == 1) v.push_back((float)c.f64());
                    else if (n5 == 2 && w5 == 2) { Reader p = c.sub(); while (!p.done()) v.push_back(p.f32()); }
                    else if (n5 == 2 && w5 == 5) v.push_back(c.f32());
                    else c.skip(w5);
                  }
                  if (!c.ok) return afail(SCANN_B200_INVALID_ARGUMENT, "malformed centre in serialized_partitioner.pb");
                  centers.push_back(std::move(v));
                } else node.skip(w4);
              }
              if (!node.ok) return afail(SCANN_B200_INVALID_ARGUMENT, "malformed k-means tree node");
            } else tree.skip(w3);
          }
        } else if (n2 == 6) {
          return afail(SCANN_B200_UNIMPLEMENTED, "two-level (bottom-up) partitioners are not supported");
        } else km.skip(w2);
      }
    } else r.skip(wt);
  }
  if (!r.ok || centers.empty()) return afail(SCANN_B200_INVALID_ARGUMENT, "serialized_partitioner.pb has no flat k-means tree");
  const size_t D = centers[0].size();
  for (auto& c : centers) if (c.size() != D) return afail(SCANN_B200_INVALID_ARGUMENT, "partitioner centres have inconsistent dimensionality");
  if (n_tokens >= 0 && (size_t)n_tokens != centers.size()) return afail(SCANN_B200_INVALID_ARGUMENT, "n_tokens %d != %zu centres", n_tokens, centers.size());
  a->n_leaves = (uint32_t)centers.size();
  a->d = (uint32_t)D;
  a->centers.resize(centers.size() * D);
  for (size_t l = 0; l < centers.size(); ++l) memcpy(&a->centers[l * D], centers[l].data(), D * 4);
  return 0;
}

// hashes/asymmetric_hashing2/training_model.cc:67-102 (Model::FromProto): CentersForAllSubspaces ->
// per block 16 GenericFeatureVectors (FLOAT or DOUBLE values).
static int parse_codebook(const std::string& bin, Assets* a) {
  Reader r(bin.data(), bin.size());
  std::vector<std::vector<std::vector<float>>> blocks;
  while (!r.done()) {
    uint64_t tag = r.varint(); int num = (int)(tag >> 3), wt = (int)(tag & 7);
    if (num == 1 && wt == 2) {
      Reader sub = r.sub();
      std::vector<std::vector<float>> cs;
      while (!sub.done()) {
        uint64_t t2 = sub.varint(); int n2 = (int)(t2 >> 3), w2 = (int)(t2 & 7);
        if (n2 == 1 && w2 == 2) {
          Reader g = sub.sub();
          std::vector<float> v;
          while (!g.done()) {
            uint64_t t3 = g.varint(); int n3 = (int)(t3 >> 3), w3 = (int)(t3 & 7);
            if (n3 == 4 && w3 == 2) { Reader p = g.sub(); while (!p.done()) v.push_back(p.f32()); }
            else if (n3 == 4 && w3 == 5) v.push_back(g.f32());
            else if (n3 == 5 && w3 == 2) { Reader p = g.sub(); while (!p.done()) v.push_back((float)p.f64()); }
            else if (n3 == 5 && w3 == 1) v.push_back((float)g.f64());
            else g.skip(w3);
          }
          if (!g.ok) return afail(SCANN_B200_INVALID_ARGUMENT, "malformed GenericFeatureVector in ah_codebook.pb");
          cs.push_back(std::move(v));
        } else sub.skip(w2);
      }
      blocks.push_back(std::move(cs));
    } else r.skip(wt);
  }
  if (!r.ok || blocks.empty()) return afail(SCANN_B200_INVALID_ARGUMENT, "ah_codebook.pb has no subspace centres");
  size_t dpb = 0;
  for (auto& b : blocks) {
    if (b.size() != 16) return afail(SCANN_B200_UNIMPLEMENTED, "only 16-centre (lut16) codebooks are supported, got %zu centres", b.size());
    for (auto& c : b) dpb = c.size() > dpb ? c.size() : dpb;
  }
  a->n_blocks = (uint32_t)blocks.size();
  a->dpb = (uint32_t)dpb;
  a->codebook.assign(blocks.size() * 16 * dpb, 0.f);
  a->block_dims.resize(blocks.size());
  for (size_t b = 0; b < blocks.size(); ++b) {
    a->block_dims[b] = (int32_t)blocks[b][0].size();
    for (size_t c = 0; c < 16; ++c) {
      if (blocks[b][c].size() != blocks[b][0].size()) return afail(SCANN_B200_INVALID_ARGUMENT, "ragged codebook block %zu", b);
      memcpy(&a->codebook[(b * 16 + c) * dpb], blocks[b][c].data(), blocks[b][c].size() * 4);
    }
  }
  return 0;
}

static std::string join_path(const std::string& dir, const std::string& p, bool* was_relative) {
  // scann_ops/cc/scann.cc:235-243: relative asset paths are re-rooted at the artifacts dir
  if (!p.empty() && p[0] == '/') { if (was_relative) *was_relative = false; return p; }
  if (was_relative) *was_relative = true;
  return dir + "/" + p;
}

static bool is_soar(const Node& cfg) {
  const Node* t = cfg.path({"partitioning", "database_spilling", "spilling_type"});
  return t && (t->scalar == "TWO_CENTER_ORTHOGONALITY_AMPLIFIED" || t->scalar == "SOAR" || t->scalar == "4");
}

int load(const char* dir, const char* assets_pbtxt, Assets** out) {
  std::unique_ptr<Assets> a(new Assets());
  std::string cfg_bin;
  if (int rc = read_file(std::string(dir) + "/scann_config.pb", &cfg_bin)) return rc;
  if (int rc = binary_to_text("ScannConfig", cfg_bin, &a->config_text)) return rc;
  std::string err;
  if (!parse_text(a->config_text, &a->config, &err)) return afail(SCANN_B200_INTERNAL, "config round trip failed: %s", err.c_str());
  Node assets;
  // scann.cc LoadArtifacts(artifacts_dir) reads <artifacts_dir>/scann_assets.pbtxt itself when it is handed no text
  std::string assets_text = assets_pbtxt ? assets_pbtxt : "";
  if (assets_text.find_first_not_of(" \t\r\n") == std::string::npos)
    if (int rc = read_file(std::string(dir) + "/scann_assets.pbtxt", &assets_text)) return rc;
  if (!parse_text(assets_text.c_str(), &assets, &err)) return afail(SCANN_B200_INVALID_ARGUMENT, "Failed to parse scann_assets.pbtxt: %s", err.c_str());
  // dependency order of scann.cc:105-233: partitioner, then tokenization, then the rest
  for (int pass = 0; pass < 3; ++pass) {
    for (const Node& as : assets.kids) {
      if (as.name != "assets" || !as.is_msg) continue;
      const Node* ty = as.find("asset_type");
      const Node* pa = as.find("asset_path");
      if (!ty || !pa) return afail(SCANN_B200_INVALID_ARGUMENT, "asset entry without type or path");
      const std::string path = join_path(dir, pa->scalar, nullptr);
      const std::string& t = ty->scalar;
      const int order = (t == "PARTITIONER" || t == "8") ? 0 : (t == "TOKENIZATION_NPY" || t == "12") ? 1 : 2;
      if (order != pass) continue;
      if (t == "PARTITIONER" || t == "8") {
        std::string bin;
        if (int rc = read_file(path, &bin)) return rc;
        if (int rc = parse_partitioner(bin, a.get())) return rc;
      } else if (t == "TOKENIZATION_NPY" || t == "12") {
        if (!a->n_leaves) return afail(SCANN_B200_INVALID_ARGUMENT, "Non-empty tokenization but no serialized partitioner is present.");
        if (int rc = read_npy(path, &a->tokens)) return rc;
        if (a->tokens.descr != "<i4") return afail(SCANN_B200_INVALID_ARGUMENT, "datapoint_to_token.npy must be int32, got %s", a->tokens.descr.c_str());
        a->has_tokens = true;
      } else if (t == "AH_CENTERS" || t == "7") {
        std::string bin;
        if (int rc = read_file(path, &bin)) return rc;
        if (int rc = parse_codebook(bin, a.get())) return rc;
      } else if (t == "AH_DATASET_NPY" || t == "11") {
        if (int rc = read_npy(path, &a->codes)) return rc;
        if (a->codes.elem != 1 || a->codes.shape.size() != 2) return afail(SCANN_B200_INVALID_ARGUMENT, "hashed_dataset.npy must be 2-D uint8");
        a->has_codes = true;
      } else if (t == "AH_DATASET_SOAR_NPY" || t == "16") {
        if (int rc = read_npy(path, &a->soar_codes)) return rc;
        if (a->soar_codes.elem != 1 || a->soar_codes.shape.size() != 2) return afail(SCANN_B200_INVALID_ARGUMENT, "hashed_dataset_soar.npy must be 2-D uint8");
        a->has_soar = true;
      } else if (t == "DATASET_NPY" || t == "9") {
        if (int rc = read_npy(path, &a->dataset)) return rc;
        if (a->dataset.descr != "<f4" || a->dataset.shape.size() != 2) return afail(SCANN_B200_INVALID_ARGUMENT, "dataset.npy must be 2-D float32");
        a->has_dataset = true;
      } else if (t == "BF16_DATASET_NPY" || t == "15") {
        if (int rc = read_npy(path, &a->bf16)) return rc;
        if (a->bf16.elem != 2 || a->bf16.shape.size() != 2) return afail(SCANN_B200_INVALID_ARGUMENT, "bfloat16_dataset.npy must be 2-D int16");
        a->has_bf16 = true;
      } else if (t == "INT8_DATASET_NPY" || t == "10") {
        if (int rc = read_npy(path, &a->int8)) return rc;
        if (a->int8.elem != 1 || a->int8.shape.size() != 2) return afail(SCANN_B200_INVALID_ARGUMENT, "int8_dataset.npy must be 2-D int8");
        a->has_int8 = true;
      } else if (t == "INT8_MULTIPLIERS_NPY" || t == "13") {
        if (int rc = read_npy(path, &a->int8_mult)) return rc;
        if (a->int8_mult.descr != "<f4" || a->int8_mult.shape.size() != 1) return afail(SCANN_B200_INVALID_ARGUMENT, "int8_multipliers.npy must be 1-D float32");
        a->has_int8_mult = true;
      } else if (t == "INT8_NORMS_NPY" || t == "14") {
        if (int rc = read_npy(path, &a->dp_norms)) return rc;
        if (a->dp_norms.descr != "<f4" || a->dp_norms.shape.size() != 1) return afail(SCANN_B200_INVALID_ARGUMENT, "dp_norms.npy must be 1-D float32");
        a->has_dp_norms = true;
      } else {
        return afail(SCANN_B200_UNIMPLEMENTED, "asset type %s is not supported by scann_b200", t.c_str());
      }
    }
  }
  if (a->has_dataset) { a->n = (uint32_t)a->dataset.shape[0]; a->d = (uint32_t)a->dataset.shape[1]; }
  else if (a->has_codes) a->n = (uint32_t)a->codes.shape[0];
  else if (a->has_bf16) { a->n = (uint32_t)a->bf16.shape[0]; a->d = (uint32_t)a->bf16.shape[1]; }
  if (a->has_int8) {
    if (!a->has_dataset && !a->has_bf16) { a->n = (uint32_t)a->int8.shape[0]; a->d = (uint32_t)a->int8.shape[1]; }
    if (!a->has_int8_mult || a->int8_mult.shape[0] != a->int8.shape[1])
      return afail(SCANN_B200_INVALID_ARGUMENT, "int8_dataset.npy needs int8_multipliers.npy with one entry per dimension");
    if (a->has_dp_norms && a->dp_norms.shape[0] == 0) a->has_dp_norms = false;  // dot product: the reference may write an empty norms file
    if (a->has_dp_norms && a->dp_norms.shape[0] != a->int8.shape[0])
      return afail(SCANN_B200_INVALID_ARGUMENT, "dp_norms.npy has %zu entries, int8_dataset.npy has %zu rows", a->dp_norms.shape[0], a->int8.shape[0]);
  }
  if (a->has_codes && a->n_blocks && a->codes.shape[1] != a->n_blocks)
    return afail(SCANN_B200_INVALID_ARGUMENT, "hashed_dataset.npy has %zu blocks, codebook has %u", a->codes.shape[1], a->n_blocks);
  if (a->has_tokens) {
    size_t want = (size_t)a->n * (is_soar(a->config) ? 2 : 1);
    size_t got = 1;
    for (size_t s : a->tokens.shape) got *= s;
    if (got != want) return afail(SCANN_B200_INVALID_ARGUMENT, "datapoint_to_token.npy has %zu entries, expected %zu", got, want);
  }
  *out = a.release();
  return 0;
}

int describe(const Assets* a, scann_b200_index_desc* d) {
  memset(d, 0, sizeof *d);
  const Node& c = a->config;
  const Node* dm = c.path({"distance_measure", "distance_measure"});
  const std::string dist = dm ? dm->scalar : "SquaredL2Distance";
  if (dist == "DotProductDistance") d->distance = SCANN_B200_DOT_PRODUCT;
  else if (dist == "SquaredL2Distance") d->distance = SCANN_B200_SQUARED_L2;
  else return afail(SCANN_B200_UNIMPLEMENTED, "distance measure %s is not supported", dist.c_str());
  d->n = a->n; d->d = a->d; d->n_leaves = a->n_leaves; d->n_blocks = a->n_blocks; d->dims_per_block = a->dpb;
  d->block_dims = a->block_dims.empty() ? nullptr : a->block_dims.data();
  d->centers = a->centers.empty() ? nullptr : a->centers.data();
  d->tokens = a->has_tokens ? reinterpret_cast<const int32_t*>(a->tokens.bytes.data() + a->tokens.data_off) : nullptr;
  d->soar = is_soar(c) ? 1 : 0;
  d->codes = a->has_codes ? reinterpret_cast<const uint8_t*>(a->codes.bytes.data() + a->codes.data_off) : nullptr;
  d->soar_codes = a->has_soar ? reinterpret_cast<const uint8_t*>(a->soar_codes.bytes.data() + a->soar_codes.data_off) : nullptr;
  d->codebook = a->codebook.empty() ? nullptr : a->codebook.data();
  d->dataset = a->has_dataset ? reinterpret_cast<const float*>(a->dataset.bytes.data() + a->dataset.data_off) : nullptr;
  d->bf16_dataset = a->has_bf16 ? reinterpret_cast<const int16_t*>(a->bf16.bytes.data() + a->bf16.data_off) : nullptr;
  d->int8_dataset = a->has_int8 ? reinterpret_cast<const int8_t*>(a->int8.bytes.data() + a->int8.data_off) : nullptr;
  d->int8_multipliers = a->has_int8_mult ? reinterpret_cast<const float*>(a->int8_mult.bytes.data() + a->int8_mult.data_off) : nullptr;
  d->dp_norms = a->has_dp_norms ? reinterpret_cast<const float*>(a->dp_norms.bytes.data() + a->dp_norms.data_off) : nullptr;
  d->overretrieve = (float)node_num(c.path({"partitioning", "database_spilling", "overretrieve_factor"}), 2.0);
  d->default_leaves = (int32_t)node_num(c.path({"partitioning", "query_spilling", "max_spill_centers"}), (double)a->n_leaves);
  d->default_final_nn = (int32_t)node_num(c.find("num_neighbors"), 1);
  d->default_pre_nn = (int32_t)node_num(c.path({"exact_reordering", "approx_num_neighbors"}), (double)d->default_final_nn);
  d->shard_world = 1;
  // partitioning { query_tokenization_type: FIXED_POINT_INT8 } (partitioning/partitioner_factory.cc:95-98)
  if (const Node* t = c.path({"partitioning", "query_tokenization_type"})) {
    if (t->scalar == "FIXED_POINT_INT8" || t->scalar == "2") d->query_tokenization_type = SCANN_B200_TOKENIZE_FIXED_POINT_INT8;
    else if (t->scalar != "FLOAT" && t->scalar != "1") return afail(SCANN_B200_UNIMPLEMENTED, "query_tokenization_type %s is not supported", t->scalar.c_str());
  }
  return 0;
}

// ---- save ---------------------------------------------------------------------------------
static std::string encode_partitioner(const scann_b200_index_desc* d) {
  // trees/kmeans_tree/kmeans_tree_node.cc:318-343: centres as packed double `dimension`, one leaf child per centre
  std::string node;
  for (uint32_t l = 0; l < d->n_leaves; ++l) {
    std::string packed;
    for (uint32_t k = 0; k < d->d; ++k) put_f64(&packed, (double)d->centers[(size_t)l * d->d + k]);
    std::string center;
    put_len(&center, 1, packed);
    put_len(&node, 1, center);
  }
  for (uint32_t l = 0; l < d->n_leaves; ++l) {
    std::string child;
    put_tag(&child, 5, 0); put_varint(&child, l);
    put_len(&node, 3, child);
  }
  std::string tree; put_len(&tree, 1, node);
  std::string km; put_len(&km, 1, tree);
  std::string sp;
  put_tag(&sp, 1, 0); put_varint(&sp, d->n_leaves);
  put_len(&sp, 2, km);
  return sp;
}
static std::string encode_codebook(const scann_b200_index_desc* d) {
  // serialization.h:28-41 DatasetSpanToCentersProto: FLOAT GenericFeatureVectors
  std::string all;
  for (uint32_t b = 0; b < d->n_blocks; ++b) {
    const uint32_t nd = d->block_dims ? (uint32_t)d->block_dims[b] : d->dims_per_block;
    std::string sub;
    for (uint32_t c = 0; c < 16; ++c) {
      std::string gfv;
      put_tag(&gfv, 1, 0); put_varint(&gfv, 2);  // feature_type = FLOAT
      std::string packed;
      for (uint32_t k = 0; k < nd; ++k) put_f32(&packed, d->codebook[((size_t)b * 16 + c) * d->dims_per_block + k]);
      put_len(&gfv, 4, packed);
      put_len(&sub, 1, gfv);
    }
    put_len(&all, 1, sub);
  }
  put_tag(&all, 2, 0); put_varint(&all, 0);  // quantization_scheme = PRODUCT
  return all;
}

int save(const char* dir, const scann_b200_index_desc* d, const char* config_text, int relative, std::string* pbtxt) {
  struct stat st;
  if (stat(dir, &st) != 0 || !S_ISDIR(st.st_mode)) return afail(SCANN_B200_FAILED_PRECONDITION, "%s is not a directory", dir);
  const std::string base(dir);
  std::string cfg_bin;
  if (int rc = text_to_binary("ScannConfig", config_text ? config_text : "", &cfg_bin)) return rc;
  if (int rc = write_file(base + "/scann_config.pb", cfg_bin.data(), cfg_bin.size())) return rc;
  std::ostringstream manifest;
  auto add = [&](const char* type, const char* file) {
    manifest << "assets {\n  asset_type: " << type << "\n  asset_path: \"" << (relative ? std::string(file) : base + "/" + file) << "\"\n}\n";
  };
  if (d->n_blocks && d->codebook) {
    const std::string b = encode_codebook(d);
    if (int rc = write_file(base + "/ah_codebook.pb", b.data(), b.size())) return rc;
    add("AH_CENTERS", "ah_codebook.pb");
  }
  if (d->n_leaves && d->centers) {
    const std::string b = encode_partitioner(d);
    if (int rc = write_file(base + "/serialized_partitioner.pb", b.data(), b.size())) return rc;
    add("PARTITIONER", "serialized_partitioner.pb");
  }
  if (d->tokens) {
    if (int rc = write_npy(base + "/datapoint_to_token.npy", "<i4", d->tokens, 4, (size_t)d->n * (d->soar ? 2 : 1), -1)) return rc;
    add("TOKENIZATION_NPY", "datapoint_to_token.npy");
  }
  if (d->codes) {
    if (int rc = write_npy(base + "/hashed_dataset.npy", "<u1", d->codes, 1, d->n, d->n_blocks)) return rc;
    add("AH_DATASET_NPY", "hashed_dataset.npy");
    if (d->soar_codes) {
      if (int rc = write_npy(base + "/hashed_dataset_soar.npy", "<u1", d->soar_codes, 1, d->n, d->n_blocks)) return rc;
      add("AH_DATASET_SOAR_NPY", "hashed_dataset_soar.npy");
    }
  }
  if (d->bf16_dataset) {
    if (int rc = write_npy(base + "/bfloat16_dataset.npy", "<i2", d->bf16_dataset, 2, d->n, d->d)) return rc;
    add("BF16_DATASET_NPY", "bfloat16_dataset.npy");
  }
  if (d->int8_dataset) {  // scann.cc:568-593 (pre_quantized_fixed_point)
    if (int rc = write_npy(base + "/int8_dataset.npy", "<i1", d->int8_dataset, 1, d->n, d->d)) return rc;
    add("INT8_DATASET_NPY", "int8_dataset.npy");
    if (d->int8_multipliers) {
      if (int rc = write_npy(base + "/int8_multipliers.npy", "<f4", d->int8_multipliers, 4, d->d, -1)) return rc;
      add("INT8_MULTIPLIERS_NPY", "int8_multipliers.npy");
    }
    if (d->dp_norms) {
      if (int rc = write_npy(base + "/dp_norms.npy", "<f4", d->dp_norms, 4, d->n, -1)) return rc;
      add("INT8_NORMS_NPY", "dp_norms.npy");
    }
  }
  if (d->dataset) {
    if (int rc = write_npy(base + "/dataset.npy", "<f4", d->dataset, 4, d->n, d->d)) return rc;
    add("DATASET_NPY", "dataset.npy");
  }
  *pbtxt = manifest.str();
  return 0;
}

}  // namespace sbassets

// ---- C ABI ----------------------------------------------------------------------------------
struct scann_b200_assets { sbassets::Assets* a; };

extern "C" {

int scann_b200_assets_load(const char* artifacts_dir, const char* assets_pbtxt, scann_b200_assets** out) {
  if (!artifacts_dir || !out) return sbassets::afail(SCANN_B200_INVALID_ARGUMENT, "null argument");
  sbassets::Assets* a = nullptr;
  if (int rc = sbassets::load(artifacts_dir, assets_pbtxt, &a)) return rc;
  *out = new scann_b200_assets{a};
  return 0;
}
void scann_b200_assets_free(scann_b200_assets* h) {
  if (!h) return;
  delete h->a;
  delete h;
}
int scann_b200_assets_describe(const scann_b200_assets* h, scann_b200_index_desc* out) {
  if (!h || !out) return sbassets::afail(SCANN_B200_INVALID_ARGUMENT, "null argument");
  return sbassets::describe(h->a, out);
}
const char* scann_b200_assets_config(const scann_b200_assets* h) { return h ? h->a->config_text.c_str() : ""; }

int scann_b200_assets_save(const char* artifacts_dir, const scann_b200_index_desc* desc, const char* config_text,
                           int relative_path, char* assets_pbtxt_out, size_t assets_pbtxt_cap) {
  if (!artifacts_dir || !desc) return sbassets::afail(SCANN_B200_INVALID_ARGUMENT, "null argument");
  std::string pbtxt;
  if (int rc = sbassets::save(artifacts_dir, desc, config_text, relative_path, &pbtxt)) return rc;
  if (assets_pbtxt_out) {
    if (pbtxt.size() + 1 > assets_pbtxt_cap) return sbassets::afail(SCANN_B200_INVALID_ARGUMENT, "assets_pbtxt buffer too small (%zu needed)", pbtxt.size() + 1);
    memcpy(assets_pbtxt_out, pbtxt.c_str(), pbtxt.size() + 1);
  }
  return 0;
}

int scann_b200_config_text_to_binary(const char* text, void* out, size_t cap, size_t* out_len) {
  std::string bin;
  if (int rc = sbassets::text_to_binary("ScannConfig", text ? text : "", &bin)) return rc;
  if (out_len) *out_len = bin.size();
  if (out) {
    if (bin.size() > cap) return sbassets::afail(SCANN_B200_INVALID_ARGUMENT, "buffer too small");
    memcpy(out, bin.data(), bin.size());
  }
  return 0;
}
int scann_b200_config_binary_to_text(const void* bin, size_t len, char* out, size_t cap) {
  std::string text;
  if (int rc = sbassets::binary_to_text("ScannConfig", std::string((const char*)bin, len), &text)) return rc;
  if (text.size() + 1 > cap) return sbassets::afail(SCANN_B200_INVALID_ARGUMENT, "buffer too small");
  memcpy(out, text.c_str(), text.size() + 1);
  return 0;
}

}  // extern "C"
