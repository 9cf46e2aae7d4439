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

void WriteBinaryFile(const std::string& path, const std::string& bytes);
std::string ReadBinaryFile(const std::string& path);

}  // namespace caffe
