// ORACLE (test infrastructure): restatement of the reference's DECOMPRESSION path, included by astc_codec.cpp.
//   physical_to_symbolic        astcenc_symbolic_physical.cpp:291-556
//   decode_ise                  astcenc_integer_sequence.cpp:651-760
//   decompress_symbolic_block   astcenc_decompress_symbolic.cpp:170-306
//   store_image_block           astcenc_image.cpp:345-563
//   astcenc_decompress_image    astcenc_entry.cpp:1274-1385 (single slice)

static inline int dec_bitrev8(int p) {   // symbolic_physical.cpp:32-38
	p = ((p & 0x0F) << 4) | ((p >> 4) & 0x0F);
	p = ((p & 0x33) << 2) | ((p >> 2) & 0x33);
	p = ((p & 0x55) << 1) | ((p >> 1) & 0x55);
	return p;
}

// read_bits :52-63. The reference reads two bytes at ptr[bitoffset >> 3]; callers pass buffers with slack.
static inline int dec_read_bits(int bitcount, int bitoffset, const uint8_t* ptr) {
	int mask = (1 << bitcount) - 1;
	ptr += bitoffset >> 3;
	bitoffset &= 7;
	int value = ptr[0] | (ptr[1] << 8);
	value >>= bitoffset;
	value &= mask;
	return value;
}

static void decode_ise(int quant_level, unsigned int character_count, const uint8_t* input_data, uint8_t* output_data, unsigned int bit_offset) {
	const ConstTables& ct = const_tables();
	uint8_t results[68];
	uint8_t tq_blocks[22];
	memset(tq_blocks, 0, sizeof(tq_blocks));
	memset(results, 0, sizeof(results));
	unsigned int bits, trits, quints;
	ise_btq(quant_level, bits, trits, quints);
	unsigned int lcounter = 0, hcounter = 0;
	for (unsigned int i = 0; i < character_count; i++) {
		results[i] = static_cast<uint8_t>(dec_read_bits((int)bits, (int)bit_offset, input_data));
		bit_offset += bits;
		if (trits) {
			static const uint8_t bits_to_read[5] = {2, 2, 1, 2, 1};
			static const uint8_t block_shift[5] = {0, 2, 4, 5, 7};
			static const uint8_t next_lcounter[5] = {1, 2, 3, 4, 0};
			static const uint8_t hcounter_incr[5] = {0, 0, 0, 0, 1};
			unsigned int tdata = (unsigned int)dec_read_bits(bits_to_read[lcounter], (int)bit_offset, input_data);
			bit_offset += bits_to_read[lcounter];
			tq_blocks[hcounter] |= tdata << block_shift[lcounter];
			hcounter += hcounter_incr[lcounter];
			lcounter = next_lcounter[lcounter];
		}
		if (quints) {
			static const uint8_t bits_to_read[3] = {3, 2, 2};
			static const uint8_t block_shift[3] = {0, 3, 5};
			static const uint8_t next_lcounter[3] = {1, 2, 0};
			static const uint8_t hcounter_incr[3] = {0, 0, 1};
			unsigned int tdata = (unsigned int)dec_read_bits(bits_to_read[lcounter], (int)bit_offset, input_data);
			bit_offset += bits_to_read[lcounter];
			tq_blocks[hcounter] |= tdata << block_shift[lcounter];
			hcounter += hcounter_incr[lcounter];
			lcounter = next_lcounter[lcounter];
		}
	}
	if (trits) {
		unsigned int trit_blocks = (character_count + 4) / 5;
		for (unsigned int i = 0; i < trit_blocks; i++) {
			const uint8_t* tritptr = ct.trits_of_integer[tq_blocks[i]];
			for (int k = 0; k < 5; k++) {
				results[5 * i + k] |= tritptr[k] << bits;
			}
		}
	}
	if (quints) {
		unsigned int quint_blocks = (character_count + 2) / 3;
		for (unsigned int i = 0; i < quint_blocks; i++) {
			const uint8_t* quintptr = ct.quints_of_integer[tq_blocks[i]];
			for (int k = 0; k < 3; k++) {
				results[3 * i + k] |= quintptr[k] << bits;
			}
		}
	}
	for (unsigned int i = 0; i < character_count; i++) {
		output_data[i] = results[i];
	}
}

static void physical_to_symbolic(const BlockSizeTables& bsd, const uint8_t pcb_in[16], SymbolicBlock& scb) {
	const ConstTables& ct = const_tables();
	uint8_t pcb[18];           // two bytes of slack for read_bits' 16-bit window
	memcpy(pcb, pcb_in, 16);
	pcb[16] = pcb[17] = 0;
	uint8_t bswapped[18];
	scb.block_type = SYM_BTYPE_NONCONST;
	int block_mode = dec_read_bits(11, 0, pcb);
	if ((block_mode & 0x1FF) == 0x1FC) {
		scb.block_type = (block_mode & 0x200) ? SYM_BTYPE_CONST_F16 : SYM_BTYPE_CONST_U16;
		scb.partition_count = 0;
		for (int i = 0; i < 4; i++) {
			scb.constant_color[i] = pcb[2 * i + 8] | (pcb[2 * i + 9] << 8);
		}
		if (bsd.dim_z > 1) {
			// 3D void extent (astcenc_symbolic_physical.cpp:348-366): six 9-bit coordinates, no reserved bits
			int v[6];
			bool ones = true;
			for (int k = 0; k < 6; k++) {
				v[k] = dec_read_bits(9, 10 + 9 * k, pcb);
				ones = ones && v[k] == 0x1FF;
			}
			if ((v[0] >= v[1] || v[2] >= v[3] || v[4] >= v[5]) && !ones) {
				scb.block_type = SYM_BTYPE_ERROR;
			}
			return;
		}
		// 2D void-extent checks
		int rsvbits = dec_read_bits(2, 10, pcb);
		if (rsvbits != 3) {
			scb.block_type = SYM_BTYPE_ERROR;
			return;
		}
		int vx_low_s = dec_read_bits(8, 12, pcb) | (dec_read_bits(5, 12 + 8, pcb) << 8);
		int vx_high_s = dec_read_bits(13, 25, pcb);
		int vx_low_t = dec_read_bits(8, 38, pcb) | (dec_read_bits(5, 38 + 8, pcb) << 8);
		int vx_high_t = dec_read_bits(13, 51, pcb);
		int all_ones = vx_low_s == 0x1FFF && vx_high_s == 0x1FFF && vx_low_t == 0x1FFF && vx_high_t == 0x1FFF;
		if ((vx_low_s >= vx_high_s || vx_low_t >= vx_high_t) && !all_ones) {
			scb.block_type = SYM_BTYPE_ERROR;
		}
		return;
	}
	unsigned int packed_index = bsd.block_mode_packed_index[block_mode];
	if (packed_index == 0xFFFF) {
		scb.block_type = SYM_BTYPE_ERROR;
		return;
	}
	const BlockMode& bm = bsd.block_modes[packed_index];
	const DecimationInfo& di = bsd.decimation_tables[bm.decimation_mode];
	int weight_count = di.weight_count;
	int weight_quant_method = bm.quant_mode;
	int is_dual_plane = bm.is_dual_plane;
	int real_weight_count = is_dual_plane ? 2 * weight_count : weight_count;
	int partition_count = dec_read_bits(2, 11, pcb) + 1;
	scb.block_mode = static_cast<uint16_t>(block_mode);
	scb.partition_count = static_cast<uint8_t>(partition_count);
	for (int i = 0; i < 16; i++) {
		bswapped[i] = static_cast<uint8_t>(dec_bitrev8(pcb[15 - i]));
	}
	bswapped[16] = bswapped[17] = 0;
	int bits_for_weights = (int)ise_sequence_bitcount((unsigned int)real_weight_count, weight_quant_method);
	int below_weights_pos = 128 - bits_for_weights;
	uint8_t indices[64];
	const WeightQuantTable& qat = ct.weight_quant[weight_quant_method];
	decode_ise(weight_quant_method, (unsigned int)real_weight_count, bswapped, indices, 0);
	if (is_dual_plane) {
		for (int i = 0; i < weight_count; i++) {
			scb.weights[i] = qat.unscramble_and_unquant_map[indices[2 * i]];
			scb.weights[i + PLANE2_OFFSET] = qat.unscramble_and_unquant_map[indices[2 * i + 1]];
		}
	} else {
		for (int i = 0; i < weight_count; i++) {
			scb.weights[i] = qat.unscramble_and_unquant_map[indices[i]];
		}
	}
	if (is_dual_plane && partition_count == 4) {
		scb.block_type = SYM_BTYPE_ERROR;
		return;
	}
	scb.color_formats_matched = 0;
	int color_formats[4] = {0, 0, 0, 0};
	int encoded_type_highpart_size = 0;
	if (partition_count == 1) {
		color_formats[0] = dec_read_bits(4, 13, pcb);
		scb.partition_index = 0;
	} else {
		encoded_type_highpart_size = (3 * partition_count) - 4;
		below_weights_pos -= encoded_type_highpart_size;
		int encoded_type = dec_read_bits(6, 13 + 10, pcb) | (dec_read_bits(encoded_type_highpart_size, below_weights_pos, pcb) << 6);
		int baseclass = encoded_type & 0x3;
		if (baseclass == 0) {
			for (int i = 0; i < partition_count; i++) {
				color_formats[i] = (encoded_type >> 2) & 0xF;
			}
			below_weights_pos += encoded_type_highpart_size;
			scb.color_formats_matched = 1;
			encoded_type_highpart_size = 0;
		} else {
			int bitpos = 2;
			baseclass--;
			for (int i = 0; i < partition_count; i++) {
				color_formats[i] = (((encoded_type >> bitpos) & 1) + baseclass) << 2;
				bitpos++;
			}
			for (int i = 0; i < partition_count; i++) {
				color_formats[i] |= (encoded_type >> bitpos) & 3;
				bitpos += 2;
			}
		}
		scb.partition_index = static_cast<uint16_t>(dec_read_bits(8, 13, pcb) | (dec_read_bits(2, 21, pcb) << 8));
		if (bsd.partitioning_packed_index[partition_count - 2][scb.partition_index] == 0xFFFF) {
			scb.block_type = SYM_BTYPE_ERROR;
			return;
		}
	}
	for (int i = 0; i < partition_count; i++) {
		scb.color_formats[i] = static_cast<uint8_t>(color_formats[i]);
	}
	int color_integer_count = 0;
	for (int i = 0; i < partition_count; i++) {
		int endpoint_class = color_formats[i] >> 2;
		color_integer_count += (endpoint_class + 1) * 2;
	}
	if (color_integer_count > 18) {
		scb.block_type = SYM_BTYPE_ERROR;
		return;
	}
	static const int color_bits_arr[5] = {-1, 115 - 4, 113 - 4 - 10, 113 - 4 - 10, 113 - 4 - 10};
	int color_bits = color_bits_arr[partition_count] - bits_for_weights - encoded_type_highpart_size;
	if (is_dual_plane) {
		color_bits -= 2;
	}
	if (color_bits < 0) {
		color_bits = 0;
	}
	int color_quant_level = ct.quant_mode_table[color_integer_count >> 1][color_bits];
	if (color_quant_level < QUANT_6) {
		scb.block_type = SYM_BTYPE_ERROR;
		return;
	}
	scb.quant_mode = static_cast<uint8_t>(color_quant_level);
	uint8_t values_to_decode[32];
	decode_ise(color_quant_level, (unsigned int)color_integer_count, pcb, values_to_decode, (partition_count == 1 ? 17 : 19 + 10));
	int valuecount_to_decode = 0;
	const uint8_t* unpack_table = ct.color_scrambled_pquant_to_uquant[scb.quant_mode - QUANT_6];
	for (int i = 0; i < partition_count; i++) {
		int vals = 2 * (color_formats[i] >> 2) + 2;
		for (int j = 0; j < vals; j++) {
			scb.color_values[i][j] = unpack_table[values_to_decode[j + valuecount_to_decode]];
		}
		valuecount_to_decode += vals;
	}
	scb.plane2_component = -1;
	if (is_dual_plane) {
		scb.plane2_component = static_cast<int8_t>(dec_read_bits(2, below_weights_pos - 2, pcb));
	}
}

static inline float error_color_nan() { return u_as_f(0xFFFFE000u); }

// decompressed texels of one block: float r,g,b,a per texel
struct DecodedBlock {
	float r[MAX_TEXELS], g[MAX_TEXELS], b[MAX_TEXELS], a[MAX_TEXELS];
};

static inline float decode_component(int v, bool lns) {   // decode_texel :66-87
	int sf = lns ? lns_to_sf16(v) : unorm16_to_sf16(v);
	return sf16_to_float((uint16_t)sf);
}

static void decompress_symbolic_block(int decode_mode, const BlockSizeTables& bsd, const SymbolicBlock& scb, bool decode_unorm8, DecodedBlock& blk) {
	unsigned int T = bsd.texel_count;
	if (scb.block_type == SYM_BTYPE_ERROR) {
		for (unsigned int i = 0; i < T; i++) {
			blk.r[i] = blk.g[i] = blk.b[i] = blk.a[i] = error_color_nan();
		}
		return;
	}
	bool u8 = decode_unorm8 || decode_mode == PRF_LDR_SRGB;
	if (scb.block_type == SYM_BTYPE_CONST_F16 || scb.block_type == SYM_BTYPE_CONST_U16) {
		float c[4];
		if (scb.block_type == SYM_BTYPE_CONST_U16) {
			for (int k = 0; k < 4; k++) {
				int ci = scb.constant_color[k];
				if (u8) {
					ci = (ci >> 8) * 257;
				}
				c[k] = sf16_to_float((uint16_t)unorm16_to_sf16(ci));
			}
		} else if (decode_mode == PRF_LDR_SRGB || decode_mode == PRF_LDR) {
			c[0] = c[1] = c[2] = c[3] = error_color_nan();
		} else {
			// float16_to_float(vint4) of the F16C builds packs with signed saturation (_mm_packs_epi32,
			// astcenc_vecmathlib_sse_4.h:1001): half-floats with the sign bit set (>= 0x8000 as int) become 0x7FFF (NaN).
			// The pinned reference is such a build, so that is the behaviour restated here.
			for (int k = 0; k < 4; k++) {
				int v = scb.constant_color[k] > 0x7FFF ? 0x7FFF : scb.constant_color[k];
				c[k] = sf16_to_float((uint16_t)v);
			}
		}
		for (unsigned int i = 0; i < T; i++) {
			blk.r[i] = c[0]; blk.g[i] = c[1]; blk.b[i] = c[2]; blk.a[i] = c[3];
		}
		return;
	}
	int partition_count = scb.partition_count;
	unsigned int packed_part = partition_count >= 2 ? bsd.partitioning_packed_index[partition_count - 2][scb.partition_index] : 0;
	const PartitionInfo& pi = bsd.partitionings[partition_count][packed_part];
	const BlockMode& bm = bsd.block_modes[bsd.block_mode_packed_index[scb.block_mode]];
	const DecimationInfo& di = bsd.decimation_tables[bm.decimation_mode];
	bool is_dual_plane = bm.is_dual_plane != 0;
	int w1[MAX_TEXELS], w2[MAX_TEXELS];
	unpack_weights(bsd, scb, di, is_dual_plane, w1, w2);
	int plane2_component = scb.plane2_component;
	for (int p = 0; p < partition_count; p++) {
		i4 ep0, ep1;
		bool rgb_lns, a_lns;
		unpack_color_endpoints(decode_mode, scb.color_formats[p], scb.color_values[p], rgb_lns, a_lns, ep0, ep1);
		int texel_count = pi.partition_texel_count[p];
		for (int j = 0; j < texel_count; j++) {
			int tix = pi.texels_of_partition[p][j];
			int wa = w1[tix], wb = is_dual_plane ? w2[tix] : w1[tix];
			i4 wv = mki4(plane2_component == 0 ? wb : wa, plane2_component == 1 ? wb : wa, plane2_component == 2 ? wb : wa, plane2_component == 3 ? wb : wa);
			i4 color = lerp_color_int(u8, ep0, ep1, wv);
			blk.r[tix] = decode_component(color.x, rgb_lns);
			blk.g[tix] = decode_component(color.y, rgb_lns);
			blk.b[tix] = decode_component(color.z, rgb_lns);
			blk.a[tix] = decode_component(color.w, a_lns);
		}
	}
}

// store_image_block: swz entries use astcenc_swz numbering (0..3 = r,g,b,a, 4 = 0, 5 = 1, 6 = Z)
static void store_image_block(void* out, int data_type, unsigned int dim_x, unsigned int dim_y, const DecodedBlock& blk, const BlockSizeTables& bsd,
                              unsigned int pos_x, unsigned int pos_y, const int swz[4], unsigned int dim_z = 1, unsigned int pos_z = 0) {
	unsigned int bx = bsd.dim_x, by = bsd.dim_y, bz = bsd.dim_z;
	bool needs_swz = swz[0] != 0 || swz[1] != 1 || swz[2] != 2 || swz[3] != 3;
	bool needs_z = swz[0] == 6 || swz[1] == 6 || swz[2] == 6 || swz[3] == 6;
	for (unsigned int z = pos_z; z < pos_z + bz && z < dim_z; z++) {
	for (unsigned int y = pos_y; y < pos_y + by && y < dim_y; y++) {
		for (unsigned int x = pos_x; x < pos_x + bx && x < dim_x; x++) {
			unsigned int idx = ((z - pos_z) * by + (y - pos_y)) * bx + (x - pos_x);
			float d[4] = {blk.r[idx], blk.g[idx], blk.b[idx], blk.a[idx]};
			size_t o = (4 * (size_t)dim_x * dim_y * z) + (4 * (size_t)dim_x * y) + 4 * (size_t)x;
			if (data_type == 0) {
				int v[7];
				v[4] = 0;
				v[5] = 255;
				for (int k = 0; k < 4; k++) {
					v[k] = f2i_rtn(clampzo(d[k]) * 255.0f);
				}
				int ov[4] = {v[0], v[1], v[2], v[3]};
				if (needs_swz) {
					v[6] = 0;
					if (needs_z) {
						float data_x = (d[0] * 2.0f) - 1.0f;
						float data_y = (d[3] * 2.0f) - 1.0f;
						float data_z = 1.0f - (data_x * data_x) - (data_y * data_y);
						data_z = maxf(data_z, 0.0f);
						data_z = (sqrtf(data_z) * 0.5f) + 0.5f;
						v[6] = f2i_rtn(minf(data_z, 1.0f) * 255.0f);
					}
					for (int k = 0; k < 4; k++) {
						ov[k] = v[swz[k]];
					}
				}
				if (d[0] != d[0]) {
					ov[0] = 0xFF; ov[1] = 0x00; ov[2] = 0xFF; ov[3] = 0xFF;
				}
				uint8_t* p = static_cast<uint8_t*>(out) + o;
				for (int k = 0; k < 4; k++) {
					p[k] = (uint8_t)ov[k];
				}
			} else {
				float v[7];
				v[4] = 0.0f;
				v[5] = 1.0f;
				v[6] = 0.0f;
				for (int k = 0; k < 4; k++) {
					v[k] = d[k];
				}
				float ov[4] = {d[0], d[1], d[2], d[3]};
				if (needs_swz) {
					if (needs_z) {
						float xN = (v[0] * 2.0f) - 1.0f;
						float yN = (v[3] * 2.0f) - 1.0f;
						float zN = 1.0f - xN * xN - yN * yN;
						if (zN < 0.0f) {
							zN = 0.0f;
						}
						v[6] = (sqrtf(zN) * 0.5f) + 0.5f;
					}
					for (int k = 0; k < 4; k++) {
						ov[k] = v[swz[k]];
					}
				}
				if (data_type == 1) {
					uint16_t* p = static_cast<uint16_t*>(out) + o;
					for (int k = 0; k < 4; k++) {
						p[k] = float_to_sf16(ov[k]);
					}
				} else {
					float* p = static_cast<float*>(out) + o;
					for (int k = 0; k < 4; k++) {
						p[k] = ov[k];
					}
				}
			}
		}
	}
	}
}

void decompress_image(const Context& ctx, const uint8_t* data, void* out, int data_type, unsigned int dim_x, unsigned int dim_y, const int swz[4], unsigned int dim_z) {
	const BlockSizeTables& bsd = *ctx.bsd;
	unsigned int bx = bsd.dim_x, by = bsd.dim_y, bz = bsd.dim_z;
	unsigned int blocks_x = (dim_x + bx - 1) / bx, blocks_y = (dim_y + by - 1) / by, blocks_z = (dim_z + bz - 1) / bz;
	for (unsigned int z = 0; z < blocks_z; z++) {
		for (unsigned int y = 0; y < blocks_y; y++) {
			for (unsigned int x = 0; x < blocks_x; x++) {
				SymbolicBlock scb;
				memset(&scb, 0, sizeof(scb));
				physical_to_symbolic(bsd, data + (((size_t)z * blocks_y + y) * blocks_x + x) * 16, scb);
				DecodedBlock blk;
				decompress_symbolic_block(ctx.config.profile, bsd, scb, data_type == 0, blk);
				store_image_block(out, data_type, dim_x, dim_y, blk, bsd, x * bx, y * by, swz, dim_z, z * bz);
			}
		}
	}
}
