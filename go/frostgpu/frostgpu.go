// Package frostgpu is the cgo binding of libfrostgpu (include/frostgpu.h).
//
// NOT BUILT IN THIS REPOSITORY'S CI: the build image has no Go toolchain (`go version` -> not
// found).  The file documents, as compilable-looking Go, the exact calls a FrostDB maintainer adds;
// the same C-ABI is exercised by frostdb_b200/_lib.py (ctypes) in the tests.
package frostgpu

/*
#cgo LDFLAGS: -lfrostgpu
#include "frostgpu.h"
#include <stdlib.h>
*/
import "C"

import (
	"errors"
	"fmt"
	"unsafe"

	"github.com/apache/arrow-go/v18/arrow"
	"github.com/apache/arrow-go/v18/arrow/cdata"
)

// Engine owns one fgpu_ctx (one GPU).
type Engine struct{ ctx *C.fgpu_ctx }

var ErrUnsupported = errors.New("frostgpu: plan not covered by the GPU path")

func lastError(rc C.int32_t) error {
	msg := C.GoString(C.fgpu_last_error())
	if rc == C.FGPU_ERR_UNSUPPORTED {
		return fmt.Errorf("%w: %s", ErrUnsupported, msg)
	}
	return fmt.Errorf("frostgpu error %d: %s", int(rc), msg)
}

// New creates the engine for one device. FGPU_ERR_NO_DEVICE is returned as an error: the caller
// keeps FrostDB's Go engine, the library never falls back to a CPU path on its own.
func New(device int) (*Engine, error) {
	cfg := C.fgpu_config{abi_version: C.FGPU_ABI_VERSION, device: C.int32_t(device)}
	var ctx *C.fgpu_ctx
	if rc := C.fgpu_init(&cfg, &ctx); rc != 0 {
		return nil, lastError(rc)
	}
	return &Engine{ctx: ctx}, nil
}

func (e *Engine) Close() { C.fgpu_shutdown(e.ctx) }

// PutPart registers one compacted Parquet part (table.go:1267 compactParts produces the bytes).
// The library copies what it needs: the slice may be reused when PutPart returns.
func (e *Engine) PutPart(table string, partID, tx uint64, parquet []byte) error {
	ct := C.CString(table)
	defer C.free(unsafe.Pointer(ct))
	rc := C.fgpu_part_put_parquet(e.ctx, ct, C.uint64_t(partID), C.uint64_t(tx),
		(*C.uint8_t)(unsafe.Pointer(&parquet[0])), C.uint64_t(len(parquet)), C.FGPU_PUT_DEFAULT)
	if rc != 0 {
		return lastError(rc)
	}
	return nil
}

// PutRecord registers an L0 part: the Arrow record Table.InsertRecord appended (parts/arrow.go:14-55,
// table.go:505-560), exported through the Arrow C Data Interface.  The library builds its column images
// from the record's buffers and releases both structs before returning.
func (e *Engine) PutRecord(table string, partID, tx uint64, rec arrow.Record) error {
	ct := C.CString(table)
	defer C.free(unsafe.Pointer(ct))
	var cs cdata.CArrowSchema
	var ca cdata.CArrowArray
	cdata.ExportArrowRecordBatch(rec, &ca, &cs)
	rc := C.fgpu_part_put_arrow(e.ctx, ct, C.uint64_t(partID), C.uint64_t(tx),
		(*C.struct_ArrowSchema)(unsafe.Pointer(&cs)), (*C.struct_ArrowArray)(unsafe.Pointer(&ca)))
	if rc != 0 {
		return lastError(rc)
	}
	return nil
}

// DropPart is called when the LSM releases the part (parts.Part.Release).
func (e *Engine) DropPart(table string, partID uint64) error {
	ct := C.CString(table)
	defer C.free(unsafe.Pointer(ct))
	if rc := C.fgpu_part_drop(e.ctx, ct, C.uint64_t(partID)); rc != 0 {
		return lastError(rc)
	}
	return nil
}

// Handle exposes the raw context to the physicalplan shim in the same module.
func (e *Engine) Handle() unsafe.Pointer { return unsafe.Pointer(e.ctx) }

// ---- multi-GPU: the exchange inside the library (fgpu_comm_*) -----------------------------------------------------
// One Engine per GPU (several in one process, or one process per GPU).  The caller moves the 128-byte handles between
// the ranks with whatever transport it has; every rank then opens the communicator with all of them in rank order.

const CommHandleBytes = C.FGPU_COMM_HANDLE_BYTES

// CommExport allocates this rank's mailbox and returns its handle.
func (e *Engine) CommExport(rank, nRanks int, slotBytes uint64) ([]byte, error) {
	h := make([]byte, CommHandleBytes)
	rc := C.fgpu_comm_export(e.ctx, C.int32_t(rank), C.int32_t(nRanks), C.uint64_t(slotBytes), (*C.uint8_t)(unsafe.Pointer(&h[0])))
	if rc != 0 {
		return nil, lastError(rc)
	}
	return h, nil
}

// CommOpen maps every rank's mailbox (handles concatenated in rank order).
func (e *Engine) CommOpen(allHandles []byte) error {
	if rc := C.fgpu_comm_open(e.ctx, (*C.uint8_t)(unsafe.Pointer(&allHandles[0]))); rc != 0 {
		return lastError(rc)
	}
	return nil
}

// CommClose releases the mailbox; call it after a barrier of the ranks.
func (e *Engine) CommClose() { C.fgpu_comm_close(e.ctx) }

// NewInProcess opens one Engine per device of this process and connects them: FrostDB is a single process, so this
// is the deployment SURVEY.md §8(b) sketches (parts are then assigned to the engine with the fewest resident bytes
// by the caller, and every engine runs the same plan through fgpu_query_execute_collective from its own goroutine).
func NewInProcess(devices []int, slotBytes uint64) ([]*Engine, error) {
	engines := make([]*Engine, 0, len(devices))
	all := make([]byte, 0, len(devices)*CommHandleBytes)
	for r, d := range devices {
		e, err := New(d)
		if err != nil {
			return nil, err
		}
		engines = append(engines, e)
		h, err := e.CommExport(r, len(devices), slotBytes)
		if err != nil {
			return nil, err
		}
		all = append(all, h...)
	}
	for _, e := range engines {
		if err := e.CommOpen(all); err != nil {
			return nil, err
		}
	}
	return engines, nil
}
