// ref_driver.cpp — a C ABI around the UNMODIFIED reference forward pass (jubruckne/Xalm, /root/reference/src), so the test suite can
// pin the oracle restatement (oracle/xalm_oracle.cpp) to the reference itself.  Test infrastructure only; built by
// oracle/Makefile.ref into oracle/_ref/libxalm_ref.so together with the reference's own infer.cpp, model.cpp, tensor.cpp,
// sampler.cpp and tokenizer.cpp, compiled where they lie (with oracle/ref_shim/ standing in for <arm_neon.h> and <print>).
// This file only CALLS the reference API (main.cpp:198-216 is the call sequence it follows).
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

#include "model.h"
#include "sampler.h"
#include "xalm.h"

namespace {
struct Ref {
	Xalm::file_info file;
	std::unique_ptr<Model> model_holder; // Model is move-only with a private constructor: keep what from_xalm returns
	std::unique_ptr<InferenceState> state;
	std::unique_ptr<Sampler> sampler;
	explicit Ref(const std::string& path) : file(Xalm::load(path)) {}
};
thread_local std::string g_err;
} // namespace

extern "C" {

const char* xref_last_error() { return g_err.c_str(); }

void* xref_load(const char* path, int context) {
	try {
		auto r = std::make_unique<Ref>(path);
		r->model_holder.reset(new Model(Model::from_xalm(r->file, context)));             // model.cpp:48 (prvalue: Model is neither copyable nor movable)
		r->state = std::make_unique<InferenceState>(r->model_holder->config);          // model.h:96
		r->sampler = std::make_unique<Sampler>(r->model_holder->config);
		return r.release();
	} catch (const std::exception& e) {
		g_err = e.what();
		return nullptr;
	}
}

void xref_free(void* h) { delete static_cast<Ref*>(h); }

// dim, hidden_dim, head_dim, n_layers, n_heads, n_kv_heads, vocab_size, max_seq_len
void xref_config(void* h, int* out) {
	const Config& c = static_cast<Ref*>(h)->model_holder->config;
	out[0] = c.dim; out[1] = c.hidden_dim; out[2] = c.head_dim; out[3] = c.n_layers;
	out[4] = c.n_heads; out[5] = c.n_kv_heads; out[6] = c.vocab_size; out[7] = c.max_seq_len;
}

// Model::forward (model.h:272); mode 0 = HYDRATE_KV_CACHE, 1 = OUTPUT_LOGITS.  logits (vocab floats) may be NULL.
int xref_forward(void* h, int token, int pos, int mode, float* logits) {
	Ref* r = static_cast<Ref*>(h);
	try {
		r->model_holder->forward(*r->state, token, pos, mode ? InferenceMode::OUTPUT_LOGITS : InferenceMode::HYDRATE_KV_CACHE);
		if (logits && mode) std::memcpy(logits, r->state->logits(), sizeof(float) * r->model_holder->config.vocab_size);
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return 1;
	}
}

int xref_sample_argmax(void* h) { // Sampler::sample_argmax (sampler.cpp:3-16) on the state's current logits
	Ref* r = static_cast<Ref*>(h);
	return r->sampler->sample_argmax(*r->state);
}

float xref_sample_prob(void* h, int index) { // Sampler::sample_prob (sampler.cpp:18-33)
	Ref* r = static_cast<Ref*>(h);
	return r->sampler->sample_prob(index, *r->state);
}

// fp16 bits of a layer's key (which = 0) or value (1) cache, (max_seq_len, n_kv_heads * head_dim)
void xref_kv(void* h, int layer, int which, uint16_t* dst, size_t n) {
	const Block& b = static_cast<Ref*>(h)->model_holder->blocks[layer];
	std::memcpy(dst, which ? b.value_cache : b.key_cache, n * sizeof(uint16_t));
}

// the residual stream x after the last forward (dim floats)
void xref_x(void* h, float* dst) {
	Ref* r = static_cast<Ref*>(h);
	std::memcpy(dst, r->state->x(), sizeof(float) * r->model_holder->config.dim);
}

unsigned long long xref_active_bytes(void* h, unsigned long long pos) { return static_cast<Ref*>(h)->model_holder->active_bytes(pos); }

} // extern "C"
