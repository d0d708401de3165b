// Package kvblockb200 is the cgo shim a maintainer of llm-d-kv-cache adds to route the kvblock hot path
// through libkvb.so (include/kvb.h).  It implements kvblock.TokenProcessor (pkg/kvcache/kvblock/
// token_processor.go:55-69) and kvblock.Index (pkg/kvcache/kvblock/index.go:120-149) one-to-one.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain (SURVEY.md, probe table).  The
// Python mirror in llm-d-kv-cache_b200/kvblock.py makes exactly the same C calls and is what the parity tests
// exercise; keep the two in lock-step.
package kvblockb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../llm-d-kv-cache_b200 -lkvb -Wl,-rpath,${SRCDIR}/../../llm-d-kv-cache_b200
#include <stdlib.h>
#include "kvb.h"
*/
import "C"

import (
	"fmt"
	"sync"
	"unsafe"

	"github.com/llm-d/llm-d-kv-cache/pkg/kvcache/kvblock"
)

func lastError(rc C.int) error {
	return fmt.Errorf("libkvb error %d: %s", int(rc), C.GoString(C.kvb_last_error()))
}

// TokenProcessor computes chained block keys on the GPU.  Same semantics as chunkedTokenDatabase:
// only full blocks, parentKey continues a chain, extraFeatures taints per block.
type TokenProcessor struct {
	cfg      kvblock.TokenProcessorConfig
	device   int
	seedHash uint64
	mu       sync.Mutex
	initHash map[string]uint64 // getInitHash(model) cache (token_processor.go:109-111)
}

var _ kvblock.TokenProcessor = (*TokenProcessor)(nil)

// NewTokenProcessor mirrors NewChunkedTokenDatabase (token_processor.go:81-106).
func NewTokenProcessor(cfg *kvblock.TokenProcessorConfig, device int) (*TokenProcessor, error) {
	if cfg == nil {
		cfg = kvblock.DefaultTokenProcessorConfig()
	}
	if cfg.BlockSize <= 0 {
		return nil, fmt.Errorf("blockSize must be greater than 0, got %d", cfg.BlockSize)
	}
	seed := []byte(cfg.HashSeed)
	var p unsafe.Pointer
	if len(seed) > 0 {
		p = unsafe.Pointer(&seed[0])
	}
	return &TokenProcessor{
		cfg: *cfg, device: device,
		seedHash: uint64(C.kvb_fnv64a(p, C.size_t(len(seed)))),
		initHash: map[string]uint64{},
	}, nil
}

func (t *TokenProcessor) BlockSize() int { return t.cfg.BlockSize }

func (t *TokenProcessor) getInitHash(model string) (uint64, error) {
	t.mu.Lock()
	defer t.mu.Unlock()
	if h, ok := t.initHash[model]; ok {
		return h, nil
	}
	cs := C.CString(model)
	defer C.free(unsafe.Pointer(cs))
	var out C.uint64_t
	if rc := C.kvb_init_hash(C.int(t.device), C.uint64_t(t.seedHash), cs, C.size_t(len(model)), &out); rc != 0 {
		return 0, lastError(rc)
	}
	t.initHash[model] = uint64(out)
	return uint64(out), nil
}

// encodeExtra is the trailing CBOR item of one block: nil -> no bytes (the kernel folds 0xf6),
// []MMHash -> array of {"Hash": text} (canonical CBOR of the struct, token_processor.go:146-148).
func encodeExtra(f *kvblock.BlockExtraFeatures) []byte {
	// The reference hands extraFeatures[i].MMHashes to the encoder, and fxamacker/cbor encodes a nil slice as null
	// (0xf6) — the same bytes as a nil *BlockExtraFeatures — not as an empty array (0x80).
	if f == nil || f.MMHashes == nil {
		return nil
	}
	out := cborHead(0x80, uint64(len(f.MMHashes)))
	for _, m := range f.MMHashes {
		out = append(out, 0xa1, 0x64, 'H', 'a', 's', 'h')
		out = append(out, cborHead(0x60, uint64(len(m.Hash)))...)
		out = append(out, m.Hash...)
	}
	return out
}

func cborHead(major byte, n uint64) []byte {
	switch {
	case n < 24:
		return []byte{major | byte(n)}
	case n < 1<<8:
		return []byte{major | 24, byte(n)}
	case n < 1<<16:
		return []byte{major | 25, byte(n >> 8), byte(n)}
	case n < 1<<32:
		return []byte{major | 26, byte(n >> 24), byte(n >> 16), byte(n >> 8), byte(n)}
	}
	b := []byte{major | 27, 0, 0, 0, 0, 0, 0, 0, 0}
	for i := 0; i < 8; i++ {
		b[1+i] = byte(n >> (56 - 8*i))
	}
	return b
}

// TokensToKVBlockKeys mirrors token_processor.go:177-205.
func (t *TokenProcessor) TokensToKVBlockKeys(
	parentKey kvblock.BlockHash, tokens []uint32, modelName string,
	extraFeatures []*kvblock.BlockExtraFeatures,
) ([]kvblock.BlockHash, error) {
	parent := uint64(parentKey)
	if parentKey == kvblock.EmptyBlockHash {
		h, err := t.getInitHash(modelName)
		if err != nil {
			return nil, err
		}
		parent = h
	}
	nblk := len(tokens) / t.cfg.BlockSize
	if nblk == 0 {
		return nil, nil
	}
	var extra []byte
	var extraOff []int64
	if extraFeatures != nil {
		if len(extraFeatures) != nblk {
			return nil, fmt.Errorf("extraFeatures length %d does not match token chunk count %d (blockSize=%d, tokens=%d)",
				len(extraFeatures), nblk, t.cfg.BlockSize, len(tokens))
		}
		extraOff = make([]int64, nblk+1)
		for i, f := range extraFeatures {
			extra = append(extra, encodeExtra(f)...)
			extraOff[i+1] = int64(len(extra))
		}
	}
	promptOff := []int64{0, int64(len(tokens))}
	keys := make([]kvblock.BlockHash, nblk) // BlockHash is uint64: written in place
	keyOff := make([]int64, 2)
	var ep *C.uint8_t
	var eo *C.int64_t
	if extraOff != nil {
		if len(extra) > 0 {
			ep = (*C.uint8_t)(unsafe.Pointer(&extra[0]))
		}
		eo = (*C.int64_t)(unsafe.Pointer(&extraOff[0]))
	}
	rc := C.kvb_hash_token_blocks(C.int(t.device),
		(*C.uint32_t)(unsafe.Pointer(&tokens[0])), (*C.int64_t)(unsafe.Pointer(&promptOff[0])),
		(*C.uint64_t)(unsafe.Pointer(&parent)), 1, C.int32_t(t.cfg.BlockSize), ep, eo,
		(*C.uint64_t)(unsafe.Pointer(&keys[0])), (*C.int64_t)(unsafe.Pointer(&keyOff[0])), nil)
	if rc != 0 {
		return nil, lastError(rc)
	}
	return keys, nil
}
