// libastcenc_b200: multi-GPU sharding and the batch scheduler (included by astcenc_b200.cu).
//
// The reference has no multi-device path; what it has is the rule that blocks are independent and that the payload is
// the row-major array of 16-byte blocks (astcenc_entry.cpp:1036: offset = ((z * yblocks + y) * xblocks + x) * 16).
// So the path shards by BLOCK ROWS with no data-path exchange, and the only collective is the gather of the payload:
//
//   slab mode  (one image, G ranks): rank g uploads image rows [r_g * block_y, r_{g+1} * block_y) only, compresses block
//              rows [r_g, r_{g+1}) with r_g = g * R / G, and sends its r * blocks_x * 16 bytes to the root, which receives
//              every slab at its offset in ONE device buffer (grouped ncclSend / ncclRecv = a gather) and copies it out once.
//   batch mode (N images, G ranks): image i belongs to rank i mod G. A rank walks its images with two device image
//              buffers: the upload of its next image runs on the copy stream under the search of the current one; the
//              payload of every finished image is sent to the root (NCCL) and from there to the caller's buffer.
//
// One process per GPU. NCCL is loaded at run time (dlopen "libnccl.so.2": inside a PyTorch process that is the copy torch
// already loaded); without it the single-GPU library works unchanged and astcenc_b200_comm_init() reports
// ASTCENC_ERR_NOT_IMPLEMENTED. The unique id travels by whatever the caller has (bench.py: torch.distributed broadcast).
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
	void* handle;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*);
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)();
	ncclResult_t (*GroupEnd)();
	const char* (*GetErrorString)(ncclResult_t);
};

static std::mutex g_nccl_mtx;
static NcclApi g_nccl;
static int g_nccl_state;      // 0 untried, 1 loaded, -1 unavailable

static bool nccl_load() {
	std::lock_guard<std::mutex> lk(g_nccl_mtx);
	if (g_nccl_state != 0) {
		return g_nccl_state > 0;
	}
	g_nccl_state = -1;
	const char* names[] = {getenv("ASTCENC_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	void* h = nullptr;
	for (const char* n : names) {
		if (n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) {
			break;
		}
	}
	if (!h) {
		return false;
	}
	g_nccl.handle = h;
	bool ok = true;
	auto sym = [&](const char* n) { void* p = dlsym(h, n); ok = ok && p != nullptr; return p; };
	g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
	g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
	g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
	g_nccl.Send = reinterpret_cast<decltype(g_nccl.Send)>(sym("ncclSend"));
	g_nccl.Recv = reinterpret_cast<decltype(g_nccl.Recv)>(sym("ncclRecv"));
	g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(sym("ncclGroupStart"));
	g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(sym("ncclGroupEnd"));
	g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
	if (!ok) {
		return false;
	}
	g_nccl_state = 1;
	return true;
}

#define NCCL_TRY(expr, onfail) do { ncclResult_t r_ = (expr); if (r_ != ncclSuccess) { \
	if (g_debug_cuda) fprintf(stderr, "astcenc_b200: %s failed: %s\n", #expr, g_nccl.GetErrorString(r_)); onfail; } } while (0)

// block rows [first, last) of rank g when `rows` block rows are cut into `world` slabs (the same rule on every rank)
static inline void slab_of(unsigned int rows, int world, int g, unsigned int& first, unsigned int& last) {
	first = (unsigned int)((size_t)rows * (size_t)g / (size_t)world);
	last = (unsigned int)((size_t)rows * (size_t)(g + 1) / (size_t)world);
}

}  // namespace

static void comm_destroy(astcenc_context* ctx) {
	if (ctx->nccl_comm && g_nccl_state > 0) {
		g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
	}
	ctx->nccl_comm = nullptr;
	ctx->rank = 0;
	ctx->world = 1;
}

extern "C" {

astcenc_error astcenc_b200_comm_unique_id(void* id_out, size_t id_bytes) {
	if (!id_out || id_bytes < sizeof(ncclUniqueId)) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	if (!nccl_load()) {
		return ASTCENC_ERR_NOT_IMPLEMENTED;
	}
	ncclUniqueId id;
	NCCL_TRY(g_nccl.GetUniqueId(&id), return ASTCENC_ERR_BAD_CONTEXT);
	memcpy(id_out, &id, sizeof(id));
	return ASTCENC_SUCCESS;
}

astcenc_error astcenc_b200_comm_init(astcenc_context* ctx, int rank, int world, const void* id_in, size_t id_bytes) {
	if (!ctx || world < 1 || rank < 0 || rank >= world) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	DeviceGuard guard(ctx->device);
	if (!guard.ok) {
		return ASTCENC_ERR_BAD_CONTEXT;
	}
	comm_destroy(ctx);
	if (world == 1) {
		return ASTCENC_SUCCESS;      // a world of one needs no communicator: the sharded calls degenerate to the local ones
	}
	if (!id_in || id_bytes < sizeof(ncclUniqueId)) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	if (!nccl_load()) {
		return ASTCENC_ERR_NOT_IMPLEMENTED;
	}
	ncclUniqueId id;
	memcpy(&id, id_in, sizeof(id));
	ncclComm_t comm = nullptr;
	NCCL_TRY(g_nccl.CommInitRank(&comm, world, id, rank), return ASTCENC_ERR_BAD_CONTEXT);
	ctx->nccl_comm = comm;
	ctx->rank = rank;
	ctx->world = world;
	return ASTCENC_SUCCESS;
}

astcenc_error astcenc_b200_comm_free(astcenc_context* ctx) {
	if (!ctx) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	DeviceGuard guard(ctx->device);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	comm_destroy(ctx);
	return ASTCENC_SUCCESS;
}

astcenc_error astcenc_b200_slab_rows(astcenc_context* ctx, unsigned int dim_y, int rank, int world, unsigned int* first_block_row, unsigned int* block_rows) {
	if (!ctx || world < 1 || rank < 0 || rank >= world || !first_block_row || !block_rows || dim_y == 0) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	unsigned int rows = (unsigned int)block_count_axis(dim_y, ctx->config.block_y);
	unsigned int a, b;
	slab_of(rows, world, rank, a, b);
	*first_block_row = a;
	*block_rows = b - a;
	return ASTCENC_SUCCESS;
}

// Gather of the slabs into the root's device buffer: root posts one receive per peer at the peer's offset, every other
// rank one send. (ncclGroupStart/End around send/recv is NCCL's gather; the root's own slab is already in place.)
static astcenc_error gather_slabs(astcenc_context* ctx, uint8_t* d_payload, unsigned int block_rows_total, size_t row_bytes, int root, cudaStream_t stream) {
	ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
	NCCL_TRY(g_nccl.GroupStart(), return ASTCENC_ERR_BAD_CONTEXT);
	if (ctx->rank == root) {
		for (int g = 0; g < ctx->world; g++) {
			if (g == root) continue;
			unsigned int a, b;
			slab_of(block_rows_total, ctx->world, g, a, b);
			if (b > a) {
				NCCL_TRY(g_nccl.Recv(d_payload + (size_t)a * row_bytes, (size_t)(b - a) * row_bytes, ncclUint8, g, comm, stream), return ASTCENC_ERR_BAD_CONTEXT);
			}
		}
	} else {
		unsigned int a, b;
		slab_of(block_rows_total, ctx->world, ctx->rank, a, b);
		if (b > a) {
			NCCL_TRY(g_nccl.Send(d_payload + (size_t)a * row_bytes, (size_t)(b - a) * row_bytes, ncclUint8, root, comm, stream), return ASTCENC_ERR_BAD_CONTEXT);
		}
	}
	NCCL_TRY(g_nccl.GroupEnd(), return ASTCENC_ERR_BAD_CONTEXT);
	return ASTCENC_SUCCESS;
}

static astcenc_error check_image_args(astcenc_context* ctx, const astcenc_image* image, const astcenc_swizzle* swizzle) {
	if (!ctx || !image || !swizzle || !image->data) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	if (ctx->config.flags & ASTCENC_FLG_DECOMPRESS_ONLY) {
		return ASTCENC_ERR_BAD_CONTEXT;
	}
	astcenc_error status = validate_compression_swizzle(*swizzle);
	if (status != ASTCENC_SUCCESS) {
		return status;
	}
	if (image->dim_x == 0 || image->dim_y == 0 || image->dim_z != 1) {
		return image->dim_z > 1 ? ASTCENC_ERR_NOT_IMPLEMENTED : ASTCENC_ERR_BAD_PARAM;
	}
	if (ctx->config.block_z > 1) {
		return ASTCENC_ERR_NOT_IMPLEMENTED;      // 3D block sizes: volumes go through astcenc_compress_image (include/astcenc.h)
	}
	if ((int)image->data_type < 0 || (int)image->data_type > 2) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	return ASTCENC_SUCCESS;
}

// ---- slab mode ----
astcenc_error astcenc_b200_compress_image_sharded(astcenc_context* ctx, astcenc_image* image, const astcenc_swizzle* swizzle, uint8_t* data_out, size_t data_len, int root) {
	astcenc_error st = check_image_args(ctx, image, swizzle);
	if (st != ASTCENC_SUCCESS) {
		return st;
	}
	if (root < 0 || root >= ctx->world || (ctx->world > 1 && !ctx->nccl_comm)) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	const DevBsd& bsd = ctx->tables->bsd;
	size_t blocks_x = block_count_axis(image->dim_x, bsd.dim_x);
	size_t blocks_y = block_count_axis(image->dim_y, bsd.dim_y);
	size_t out_bytes = blocks_x * blocks_y * 16;
	if (ctx->rank == root && (!data_out || data_len < out_bytes)) {
		return data_out ? ASTCENC_ERR_OUT_OF_MEM : ASTCENC_ERR_BAD_PARAM;
	}
	if (ctx->config.a_scale_radius != 0 && ctx->world > 1) {
		return ASTCENC_ERR_NOT_IMPLEMENTED;      // the alpha pre-pass filters across slab borders: it would need the neighbours' rows
	}
	DeviceGuard guard(ctx->device);
	if (!guard.ok) {
		return ASTCENC_ERR_BAD_CONTEXT;
	}
	size_t bpt = image->data_type == ASTCENC_TYPE_U8 ? 4 : image->data_type == ASTCENC_TYPE_F16 ? 8 : 16;
	size_t row_bytes = (size_t)image->dim_x * bpt;
	st = ensure_buffer(ctx->d_image, ctx->d_image_bytes, row_bytes * image->dim_y);
	if (st == ASTCENC_SUCCESS) {
		st = ensure_buffer(ctx->d_out, ctx->d_out_bytes, out_bytes);
	}
	if (st != ASTCENC_SUCCESS) {
		return st;
	}
	unsigned int r0, r1;
	slab_of((unsigned int)blocks_y, ctx->world, ctx->rank, r0, r1);
	int swz[4] = {(int)swizzle->r, (int)swizzle->g, (int)swizzle->b, (int)swizzle->a};
	ctx->last_h2d = ctx->last_d2h = 0;
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
	if (r1 > r0) {
		// only this rank's rows travel: the bands of the upload plan are cut inside [r0, r1)
		UploadPlan up;
		up.host = static_cast<const uint8_t*>(image->data[0]);
		up.device = ctx->d_image;
		up.row_bytes = row_bytes;
		up.bands = ctx->knobs.upload_bands;
		size_t y0 = (size_t)r0 * bsd.dim_y, y1 = (size_t)r1 * bsd.dim_y;
		if (y1 > image->dim_y) y1 = image->dim_y;
		ctx->last_h2d = (y1 - y0) * row_bytes;
		st = launch_slab(ctx, ctx->d_image, (int)image->data_type, image->dim_x, image->dim_y, swz, r0, r1 - r0, ctx->d_out + (size_t)r0 * blocks_x * 16, ctx->stream, &up);
		if (st != ASTCENC_SUCCESS) {
			return st;
		}
	}
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
	CUDA_TRY(cudaEventRecord(ctx->ev_g0, ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
	if (ctx->world > 1) {
		st = gather_slabs(ctx, ctx->d_out, (unsigned int)blocks_y, blocks_x * 16, root, ctx->stream);
		if (st != ASTCENC_SUCCESS) {
			return st;
		}
	}
	CUDA_TRY(cudaEventRecord(ctx->ev_g1, ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
	if (ctx->rank == root) {
		CUDA_TRY(cudaMemcpyAsync(data_out, ctx->d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
		ctx->last_d2h = out_bytes;
	}
	CUDA_TRY(cudaStreamSynchronize(ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
	cudaEventElapsedTime(&ctx->last_compress_ms, ctx->ev0, ctx->ev1);
	cudaEventElapsedTime(&ctx->last_gather_ms, ctx->ev_g0, ctx->ev_g1);
	ctx->last_kernel_ms = ctx->last_compress_ms;
	return ASTCENC_SUCCESS;
}

// ---- batch mode ----
// images[i] / data_out[i]: image i of the batch and where its payload goes. Rank g works on the images with i mod world == g
// (entries of other ranks are not touched and may be NULL); payloads end up on the root only. All images of a batch share
// dimensions and data type (the batch scheduler of BASELINE.json configs[4]: 64 x 4096^2, 8 per GPU).
astcenc_error astcenc_b200_compress_batch(astcenc_context* ctx, astcenc_image* const* images, unsigned int image_count, const astcenc_swizzle* swizzle,
                                          uint8_t* const* data_out, size_t data_len_each, int root) {
	if (!ctx || !images || image_count == 0 || !swizzle) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	if (root < 0 || root >= ctx->world || (ctx->world > 1 && !ctx->nccl_comm)) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	// the first image this rank owns fixes the geometry
	const astcenc_image* ref = nullptr;
	for (unsigned int i = (unsigned int)ctx->rank; i < image_count; i += (unsigned int)ctx->world) {
		astcenc_error st = check_image_args(ctx, images[i], swizzle);
		if (st != ASTCENC_SUCCESS) {
			return st;
		}
		if (!ref) {
			ref = images[i];
		} else if (images[i]->dim_x != ref->dim_x || images[i]->dim_y != ref->dim_y || images[i]->data_type != ref->data_type) {
			return ASTCENC_ERR_BAD_PARAM;
		}
	}
	// (a rank without images still takes part in the gather when it is the root: it needs the geometry from the caller)
	if (!ref) {
		for (unsigned int i = 0; i < image_count && !ref; i++) {
			if (images[i]) ref = images[i];
		}
		if (!ref) {
			return ASTCENC_ERR_BAD_PARAM;
		}
	}
	const DevBsd& bsd = ctx->tables->bsd;
	size_t blocks_x = block_count_axis(ref->dim_x, bsd.dim_x);
	size_t blocks_y = block_count_axis(ref->dim_y, bsd.dim_y);
	size_t out_bytes = blocks_x * blocks_y * 16;
	if (ctx->rank == root) {
		if (!data_out || data_len_each < out_bytes) {
			return data_out ? ASTCENC_ERR_OUT_OF_MEM : ASTCENC_ERR_BAD_PARAM;
		}
		for (unsigned int i = 0; i < image_count; i++) {
			if (!data_out[i]) return ASTCENC_ERR_BAD_PARAM;
		}
	}
	DeviceGuard guard(ctx->device);
	if (!guard.ok) {
		return ASTCENC_ERR_BAD_CONTEXT;
	}
	size_t bpt = ref->data_type == ASTCENC_TYPE_U8 ? 4 : ref->data_type == ASTCENC_TYPE_F16 ? 8 : 16;
	size_t slice_bytes = (size_t)ref->dim_x * ref->dim_y * bpt;
	astcenc_error st = ensure_buffer(ctx->d_image, ctx->d_image_bytes, slice_bytes);
	if (st == ASTCENC_SUCCESS) st = ensure_buffer(ctx->d_image2, ctx->d_image2_bytes, slice_bytes);
	// one payload slot per round of the batch on the root (it receives world payloads per round), two rounds in flight
	unsigned int rounds = (image_count + (unsigned int)ctx->world - 1) / (unsigned int)ctx->world;
	size_t slots = ctx->rank == root ? (size_t)ctx->world * 2 : 2;
	if (st == ASTCENC_SUCCESS) st = ensure_buffer(ctx->d_out, ctx->d_out_bytes, out_bytes * slots);
	if (st != ASTCENC_SUCCESS) {
		return st;
	}
	int swz[4] = {(int)swizzle->r, (int)swizzle->g, (int)swizzle->b, (int)swizzle->a};
	ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
	ctx->last_h2d = ctx->last_d2h = 0;
	uint8_t* bufs[2] = {ctx->d_image, ctx->d_image2};
	// three streams: uploads (copy_stream), search + gather (stream), payload copies to the host (down_stream); per parity
	// of the round one event for "image uploaded", "image buffer free", "payload slots filled", "payload slots copied out"
	struct Events {
		cudaEvent_t e[8];
		cudaStream_t down;
		Events() : down(nullptr) { for (auto& x : e) x = nullptr; }
		~Events() {
			for (auto& x : e) if (x) cudaEventDestroy(x);
			if (down) cudaStreamDestroy(down);
		}
	} ev;
	for (auto& x : ev.e) {
		CUDA_TRY(cudaEventCreateWithFlags(&x, cudaEventDisableTiming), return ASTCENC_ERR_BAD_CONTEXT);
	}
	CUDA_TRY(cudaStreamCreateWithFlags(&ev.down, cudaStreamNonBlocking), return ASTCENC_ERR_BAD_CONTEXT);
	cudaEvent_t* up_done = ev.e;
	cudaEvent_t* buf_free = ev.e + 2;
	cudaEvent_t* slots_full = ev.e + 4;
	cudaEvent_t* slots_free = ev.e + 6;
	CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
	auto upload = [&](unsigned int round) -> astcenc_error {
		unsigned int i = round * (unsigned int)ctx->world + (unsigned int)ctx->rank;
		if (i >= image_count) {
			return ASTCENC_SUCCESS;
		}
		int b = (int)(round & 1);
		if (round >= 2) {
			CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, buf_free[b], 0), return ASTCENC_ERR_BAD_CONTEXT);
		}
		CUDA_TRY(cudaMemcpyAsync(bufs[b], images[i]->data[0], slice_bytes, cudaMemcpyHostToDevice, ctx->copy_stream), return ASTCENC_ERR_BAD_CONTEXT);
		CUDA_TRY(cudaEventRecord(up_done[b], ctx->copy_stream), return ASTCENC_ERR_BAD_CONTEXT);
		ctx->last_h2d += slice_bytes;
		return ASTCENC_SUCCESS;
	};
	{
		// a pass of an earlier call may still read the image buffers
		std::lock_guard<std::mutex> lk(ctx->launch_mtx);
		if (ctx->scratch_used) {
			CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->scratch_done, 0), return ASTCENC_ERR_BAD_CONTEXT);
		}
	}
	// (round 0 has nothing to hide its upload behind: it goes up in bands under its own first wave, as in astcenc_compress_image)
	for (unsigned int round = 0; round < rounds; round++) {
		unsigned int i = round * (unsigned int)ctx->world + (unsigned int)ctx->rank;
		int b = (int)(round & 1);
		// payload slots of this round: the root keeps one per rank, the others one; reused every second round
		uint8_t* slot0 = ctx->d_out + (size_t)b * out_bytes * (ctx->rank == root ? (size_t)ctx->world : 1);
		uint8_t* mine = ctx->rank == root ? slot0 + (size_t)ctx->rank * out_bytes : slot0;
		if (round >= 2 && ctx->rank == root) {
			CUDA_TRY(cudaStreamWaitEvent(ctx->stream, slots_free[b], 0), return ASTCENC_ERR_BAD_CONTEXT);
		}
		if (i < image_count) {
			UploadPlan first;
			first.host = static_cast<const uint8_t*>(images[i]->data[0]);
			first.device = bufs[b];
			first.row_bytes = (size_t)ref->dim_x * bpt;
			first.bands = ctx->knobs.upload_bands;
			if (round != 0) {
				CUDA_TRY(cudaStreamWaitEvent(ctx->stream, up_done[b], 0), return ASTCENC_ERR_BAD_CONTEXT);
			} else {
				ctx->last_h2d += slice_bytes;
			}
			st = launch_slab(ctx, bufs[b], (int)ref->data_type, ref->dim_x, ref->dim_y, swz, 0, (unsigned int)blocks_y, mine, ctx->stream, round == 0 ? &first : nullptr);
			if (st != ASTCENC_SUCCESS) {
				return st;
			}
			CUDA_TRY(cudaEventRecord(buf_free[b], ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
		}
		// the next image goes up while this one is searched (enqueued AFTER the search: a copy from pageable memory blocks
		// the host thread for its duration, and the search must already be running then)
		if (round + 1 < rounds) {
			st = upload(round + 1);
			if (st != ASTCENC_SUCCESS) {
				return st;
			}
		}
		if (ctx->world > 1) {
			NCCL_TRY(g_nccl.GroupStart(), return ASTCENC_ERR_BAD_CONTEXT);
			if (ctx->rank == root) {
				for (int g = 0; g < ctx->world; g++) {
					unsigned int gi = round * (unsigned int)ctx->world + (unsigned int)g;
					if (g != root && gi < image_count) {
						NCCL_TRY(g_nccl.Recv(slot0 + (size_t)g * out_bytes, out_bytes, ncclUint8, g, comm, ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
					}
				}
			} else if (i < image_count) {
				NCCL_TRY(g_nccl.Send(mine, out_bytes, ncclUint8, root, comm, ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
			}
			NCCL_TRY(g_nccl.GroupEnd(), return ASTCENC_ERR_BAD_CONTEXT);
		}
		if (ctx->rank == root) {
			CUDA_TRY(cudaEventRecord(slots_full[b], ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
			CUDA_TRY(cudaStreamWaitEvent(ev.down, slots_full[b], 0), return ASTCENC_ERR_BAD_CONTEXT);
			for (int g = 0; g < ctx->world; g++) {
				unsigned int gi = round * (unsigned int)ctx->world + (unsigned int)g;
				if (gi < image_count) {
					CUDA_TRY(cudaMemcpyAsync(data_out[gi], slot0 + (size_t)g * out_bytes, out_bytes, cudaMemcpyDeviceToHost, ev.down), return ASTCENC_ERR_BAD_CONTEXT);
					ctx->last_d2h += out_bytes;
				}
			}
			CUDA_TRY(cudaEventRecord(slots_free[b], ev.down), return ASTCENC_ERR_BAD_CONTEXT);
		}
	}
	CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
	CUDA_TRY(cudaStreamSynchronize(ctx->stream), return ASTCENC_ERR_BAD_CONTEXT);
	CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream), return ASTCENC_ERR_BAD_CONTEXT);
	CUDA_TRY(cudaStreamSynchronize(ev.down), return ASTCENC_ERR_BAD_CONTEXT);
	cudaEventElapsedTime(&ctx->last_compress_ms, ctx->ev0, ctx->ev1);
	ctx->last_kernel_ms = ctx->last_compress_ms;
	ctx->last_gather_ms = 0.0f;
	return ASTCENC_SUCCESS;
}

astcenc_error astcenc_b200_comm_last_timing(astcenc_context* ctx, float* compress_ms, float* gather_ms) {
	if (!ctx) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	if (compress_ms) *compress_ms = ctx->last_compress_ms;
	if (gather_ms) *gather_ms = ctx->last_gather_ms;
	return ASTCENC_SUCCESS;
}

}  // extern "C"
