// GPUScan: the ScanPhysicalPlan (query/physicalplan/physicalplan.go:32-35) that replaces the planned
// prefix TableScan -> PredicateFilter -> Projection -> HashAggregate(partial) -> Synchronizer ->
// HashAggregate(final) (physicalplan.go:333-474) by one call into libfrostgpu.
//
// NOT BUILT HERE (no Go toolchain in the image); it belongs in FrostDB's query/physicalplan package.
package physicalplan

/*
#include "frostgpu.h"
*/
import "C"

import (
	"context"
	"unsafe"

	"github.com/apache/arrow-go/v18/arrow/cdata"
	"github.com/apache/arrow-go/v18/arrow/memory"

	"github.com/polarsignals/frostdb/query/logicalplan"
)

// WithGPUEngine is the new Option next to WithOverrideInput (physicalplan.go:279-285).
func WithGPUEngine(ctx unsafe.Pointer) Option {
	return func(o *execOptions) { o.gpu = (*C.fgpu_ctx)(ctx) }
}

type GPUScan struct {
	gpu   *C.fgpu_ctx
	table logicalplan.TableReader
	plan  *cPlan // POD mirror of the optimised logical plan (see INTEGRATION.md, "Plan descriptor")
	next  PhysicalPlan
}

func (s *GPUScan) SetNext(p PhysicalPlan) { s.next = p }

func (s *GPUScan) Draw() *Diagram {
	return &Diagram{Details: "GPUScan (" + s.plan.describe() + ")", Child: s.next.Draw()}
}

// Execute mirrors TableScan.Execute (physicalplan.go:114-166): take the read transaction, run the
// fused prefix on the GPU, push every result record into the next operator, then Finish it.
func (s *GPUScan) Execute(ctx context.Context, _ memory.Allocator) error {
	return s.table.View(ctx, func(ctx context.Context, tx uint64) error {
		var q *C.fgpu_query
		if rc := C.fgpu_query_prepare(s.gpu, s.plan.c(), &q); rc != 0 {
			return gpuError(rc)
		}
		defer C.fgpu_query_free(q)
		var res *C.fgpu_result
		if rc := C.fgpu_query_execute(s.gpu, q, C.uint64_t(tx), &res); rc != 0 {
			return gpuError(rc)
		}
		defer C.fgpu_result_free(res)
		for {
			var cs cdata.CArrowSchema
			var ca cdata.CArrowArray
			rc := C.fgpu_result_next(res, (*C.struct_ArrowSchema)(unsafe.Pointer(&cs)), (*C.struct_ArrowArray)(unsafe.Pointer(&ca)))
			if rc == C.FGPU_ERR_END {
				break
			}
			if rc != 0 {
				return gpuError(rc)
			}
			rec, err := cdata.ImportCRecordBatch(&ca, &cs) // zero copy; buffers are freed by the release callbacks
			if err != nil {
				return err
			}
			err = s.next.Callback(ctx, rec)
			rec.Release()
			if err != nil {
				return err
			}
		}
		return s.next.Finish(ctx)
	})
}
