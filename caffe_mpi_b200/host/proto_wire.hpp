// proto_wire.hpp -- protobuf BINARY wire format for the two files Solver::Snapshot writes: the
// .caffemodel (NetParameter with every layer's blobs) and the .solverstate (SolverState: iter, learned_net, history,
// current_step).  Hand-written (there is no libprotobuf in the toolchain); tests/test_snapshot_cpu.py checks both
// directions against google.protobuf built from the same message schema.
//
// Reference map: src/caffe/proto/caffe.proto:15-35 (BlobShape, BlobProto), :88-146 (NetParameter: name = 1, layer = 100,
// V1 layers = 2), :303-308 (SolverState), LayerParameter name = 1 / type = 2 / bottom = 3 / top = 4 / blobs = 7;
// Blob::ToProto / FromProto (src/caffe/blob.cpp:352-476): NVCaffe writes shape + raw_data_type + raw_data, and reads
// data / double_data / raw_data (FLOAT, FLOAT16, DOUBLE) and the legacy num/channels/height/width dims;
// Net::ToProto, Net::CopyTrainedLayersFrom (net.cpp), SGDSolver::SnapshotSolverStateToBinaryProto / RestoreSolverState.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace caffe {

struct BlobData { std::vector<int> shape; std::vector<float> data; };
struct LayerWeights { std::string name, type; std::vector<std::string> bottom, top; std::vector<BlobData> blobs; };
struct NetWeights { std::string name; std::vector<LayerWeights> layers; };
struct SolverStateData { int iter = 0, current_step = 0; std::string learned_net; std::vector<BlobData> history; };

// raw_format = true: NVCaffe's BlobProto (shape, raw_data_type = FLOAT, raw_data bytes); false: BVLC's (shape, packed data)
std::string SerializeNetWeights(const NetWeights& net, bool raw_format = true);
NetWeights ParseNetWeights(const std::string& bytes);                 // throws FatalError on malformed input
std::string SerializeSolverState(const SolverStateData& st, bool raw_format = true);
SolverStateData ParseSolverState(const std::string& bytes);

// Datum (caffe.proto:43-56), the record type of Caffe's LMDB / LevelDB image databases.  ParseDatum is the zero-copy form of
// Datum::ParseFromArray: `data` is a view into the parsed buffer (for an LMDB value: into the file mapping), so a parser thread
// copies the pixel bytes once, straight into the pinned batch buffer.  Returns false on malformed input, like ParseFromArray.
struct Datum {
  int channels = 0, height = 0, width = 0, label = 0;
  const uint8_t* data = nullptr;          // field 4 (bytes), not owned
  size_t data_size = 0;
  std::vector<float> float_data;          // field 6, packed or unpacked
  bool encoded = false;                   // field 7: `data` holds a compressed image (JPEG / PNG) -- not decodable here
  uint32_t record_id = 0;                 // field 8: assigned by the reader (data_reader.cpp:224)
};
bool ParseDatum(const void* bytes, size_t n, Datum* d);
std::string SerializeDatum(int channels, int height, int width, const void* data, size_t data_size, int label, bool encoded = false,
                           const std::vector<float>* float_data = nullptr);
// a bare BlobProto file: transform_param.mean_file (data_transformer.cpp:31-38 reads it with ReadProtoFromBinaryFileOrDie)
BlobData ParseBlobProto(const std::string& bytes);
std::string SerializeBlobProto(const BlobData& b, bool raw_format = false);

void WriteBinaryFile(const std::string& path, const std::string& bytes);
std::string ReadBinaryFile(const std::string& path);

}  // namespace caffe
