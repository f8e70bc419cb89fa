// GPUScan: the ScanPhysicalPlan (query/physicalplan/physicalplan.go:32-35) that replaces the planned prefix
//
//	TableScan -> PredicateFilter -> Projection -> HashAggregate(partial) -> Synchronizer -> HashAggregate(final)
//
// (physicalplan.go:333-474) by one call into libfrostgpu.  This file belongs in FrostDB's query/physicalplan
// package, next to physicalplan.go; together with the two small patches quoted at the bottom it is everything the
// Go side needs.
//
// NOT BUILT HERE: the build image has no Go toolchain (`go version` -> not found), so this file has never seen a
// compiler.  It is written against the reference at 906ebbae (types and fields cited inline) and against
// include/frostgpu.h; the same C-ABI is exercised through ctypes by frostdb_b200/_lib.py in the tests.
package physicalplan

/*
#cgo LDFLAGS: -lfrostgpu
#include <stdlib.h>
#include "frostgpu.h"

// cgo cannot take the address of a Go function: the regex leaf's matcher is this C trampoline around the exported
// Go function below (//export frostgpuRegexMatch).
extern int32_t frostgpuRegexMatch(void* user, uint8_t* value, uint64_t len);
static inline fgpu_match_fn frostgpu_regex_trampoline(void) { return (fgpu_match_fn)frostgpuRegexMatch; }
*/
import "C"

import (
	"context"
	"errors"
	"fmt"
	"regexp"
	"runtime/cgo"
	"strings"
	"unsafe"

	"github.com/apache/arrow-go/v18/arrow/cdata"
	"github.com/apache/arrow-go/v18/arrow/memory"
	"github.com/apache/arrow-go/v18/arrow/scalar"

	"github.com/polarsignals/frostdb/query/logicalplan"
)

// ErrGPUUnsupported makes Build keep the Go operator chain for this plan (never a fallback inside the library).
var ErrGPUUnsupported = errors.New("plan not covered by the GPU path")

func gpuError(rc C.int32_t) error {
	msg := C.GoString(C.fgpu_last_error())
	if rc == C.FGPU_ERR_UNSUPPORTED {
		return fmt.Errorf("%w: %s", ErrGPUUnsupported, msg)
	}
	return fmt.Errorf("frostgpu error %d: %s", int(rc), msg)
}

// WithGPUEngine is the new Option next to WithOverrideInput (physicalplan.go:279-285).  `ctx` is
// frostgpu.Engine.Handle().  Patch 1 (physicalplan.go:259-263) adds the field:
//
//	type execOptions struct {
//		orderedAggregations bool
//		overrideInput       []PhysicalPlan
//		readMode            logicalplan.ReadMode
//		gpu                 unsafe.Pointer // *C.fgpu_ctx, nil = Go operators only
//	}
func WithGPUEngine(ctx unsafe.Pointer) Option {
	return func(o *execOptions) { o.gpu = ctx }
}

//export frostgpuRegexMatch
func frostgpuRegexMatch(user unsafe.Pointer, value *C.uint8_t, n C.uint64_t) C.int32_t {
	re := cgo.Handle(uintptr(user)).Value().(*regexp.Regexp) // compiled like filter.go:104-123 compiles it
	if re.Match(unsafe.Slice((*byte)(unsafe.Pointer(value)), int(n))) {
		return 1
	}
	return 0
}

// cPlan is the POD mirror (fgpu_plan) of the optimised logical plan prefix, allocated in C memory so that no Go
// pointer is retained by C.  Children precede parents in exprs, as include/frostgpu.h asks.
type cPlan struct {
	plan    C.fgpu_plan
	exprs   []C.fgpu_expr // backing store copied into C memory by finish()
	groups  []C.int32_t
	aggs    []C.fgpu_agg
	cstrs   []unsafe.Pointer // C strings / byte buffers to free
	handles []cgo.Handle     // regex handles to delete
	text    string           // for Draw()
}

func (p *cPlan) cstring(s string) *C.char {
	c := C.CString(s)
	p.cstrs = append(p.cstrs, unsafe.Pointer(c))
	return c
}

// add flattens one logicalplan.Expr (expr.go) and returns its index.
func (p *cPlan) add(e logicalplan.Expr) (C.int32_t, error) {
	var x C.fgpu_expr
	x.left, x.right = -1, -1
	switch e := e.(type) {
	case *logicalplan.AliasExpr: // expr.go:1000-1003: the alias is a host-side name
		return p.add(e.Expr)
	case *logicalplan.Column: // expr.go:292-294
		x.kind = C.FGPU_EXPR_COLUMN
		x.name = p.cstring(e.ColumnName)
	case *logicalplan.DynamicColumn: // expr.go:518-520
		x.kind = C.FGPU_EXPR_DYNCOLUMN
		x.name = p.cstring(e.ColumnName)
	case *logicalplan.LiteralExpr: // expr.go:586-588
		x.kind = C.FGPU_EXPR_LITERAL
		switch v := e.Value.(type) {
		case *scalar.Null:
			x.literal._type = C.FGPU_SCALAR_NULL
		case *scalar.Int64:
			x.literal._type = C.FGPU_SCALAR_INT64
			x.literal.i64 = C.int64_t(v.Value)
		case *scalar.Float64:
			x.literal._type = C.FGPU_SCALAR_FLOAT64
			x.literal.f64 = C.double(v.Value)
		case *scalar.String:
			b := v.Value.Bytes()
			x.literal._type = C.FGPU_SCALAR_STRING
			x.literal.bytes = (*C.uint8_t)(C.CBytes(b))
			x.literal.len = C.uint64_t(len(b))
			p.cstrs = append(p.cstrs, unsafe.Pointer(x.literal.bytes))
		case *scalar.Binary:
			b := v.Value.Bytes()
			x.literal._type = C.FGPU_SCALAR_STRING
			x.literal.bytes = (*C.uint8_t)(C.CBytes(b))
			x.literal.len = C.uint64_t(len(b))
			p.cstrs = append(p.cstrs, unsafe.Pointer(x.literal.bytes))
		default:
			return -1, fmt.Errorf("%w: literal %s", ErrGPUUnsupported, e.Value.DataType())
		}
	case *logicalplan.BinaryExpr: // expr.go:105-109; Op values are the wire values (expr.go:13-33)
		l, err := p.add(e.Left)
		if err != nil {
			return -1, err
		}
		r, err := p.add(e.Right)
		if err != nil {
			return -1, err
		}
		x.kind = C.FGPU_EXPR_BINARY
		x.op = C.int32_t(e.Op)
		x.left, x.right = l, r
		if e.Op == logicalplan.OpRegexMatch || e.Op == logicalplan.OpRegexNotMatch {
			lit, ok := e.Right.(*logicalplan.LiteralExpr)
			if !ok {
				return -1, fmt.Errorf("%w: regex against a non-literal", ErrGPUUnsupported)
			}
			re, err := regexp.Compile(string(lit.Value.(*scalar.String).Value.Bytes())) // filter.go:104-123
			if err != nil {
				return -1, err
			}
			h := cgo.NewHandle(re)
			p.handles = append(p.handles, h)
			x.match = C.frostgpu_regex_trampoline()
			x.match_user = unsafe.Pointer(uintptr(h)) //nolint:govet // an opaque cookie, never dereferenced by C
		}
	default:
		return -1, fmt.Errorf("%w: expression %T", ErrGPUUnsupported, e)
	}
	p.exprs = append(p.exprs, x)
	return C.int32_t(len(p.exprs) - 1), nil
}

// finish copies the Go-side slices into C memory and wires the fgpu_plan.
func (p *cPlan) finish(table string, kind C.int32_t, filter C.int32_t) {
	copyTo := func(src unsafe.Pointer, n int, size uintptr) unsafe.Pointer {
		if n == 0 {
			return nil
		}
		dst := C.malloc(C.size_t(uintptr(n) * size))
		copy(unsafe.Slice((*byte)(dst), uintptr(n)*size), unsafe.Slice((*byte)(src), uintptr(n)*size))
		p.cstrs = append(p.cstrs, dst)
		return dst
	}
	p.plan.table = p.cstring(table)
	p.plan.kind = kind
	p.plan.filter = filter
	p.plan.n_exprs = C.int32_t(len(p.exprs))
	if len(p.exprs) > 0 {
		p.plan.exprs = (*C.fgpu_expr)(copyTo(unsafe.Pointer(&p.exprs[0]), len(p.exprs), unsafe.Sizeof(p.exprs[0])))
	}
	p.plan.n_group_by = C.int32_t(len(p.groups))
	if len(p.groups) > 0 {
		p.plan.group_by = (*C.int32_t)(copyTo(unsafe.Pointer(&p.groups[0]), len(p.groups), unsafe.Sizeof(p.groups[0])))
	}
	p.plan.n_aggs = C.int32_t(len(p.aggs))
	if len(p.aggs) > 0 {
		p.plan.aggs = (*C.fgpu_agg)(copyTo(unsafe.Pointer(&p.aggs[0]), len(p.aggs), unsafe.Sizeof(p.aggs[0])))
	}
}

func (p *cPlan) free() {
	for _, c := range p.cstrs {
		C.free(c)
	}
	for _, h := range p.handles {
		h.Delete()
	}
	p.cstrs, p.handles = nil, nil
}

// encodeGPUPlan covers TableScan [-> Filter]* [-> Projection(pass-through / + - * / [as alias])] -> Aggregation |
// Distinct; anything else returns ErrGPUUnsupported and Build keeps the Go operators.  `nodes` is the plan in
// scan-first order (the reverse of LogicalPlan.Input links, logicalplan.go:17-29); the returned int is the number
// of nodes the GPUScan consumes.
func encodeGPUPlan(nodes []*logicalplan.LogicalPlan) (*cPlan, int, error) {
	if len(nodes) == 0 || nodes[0].TableScan == nil {
		return nil, 0, ErrGPUUnsupported
	}
	scan := nodes[0].TableScan
	p := &cPlan{}
	ok := false
	defer func() {
		if !ok {
			p.free()
		}
	}()
	filter := C.int32_t(-1)
	andWith := func(e logicalplan.Expr) error {
		i, err := p.add(e)
		if err != nil {
			return err
		}
		if filter < 0 {
			filter = i
			return nil
		}
		var x C.fgpu_expr
		x.kind, x.op, x.left, x.right = C.FGPU_EXPR_BINARY, C.int32_t(logicalplan.OpAnd), filter, i
		p.exprs = append(p.exprs, x)
		filter = C.int32_t(len(p.exprs) - 1)
		return nil
	}
	if scan.Filter != nil { // FilterPushDown left it here (optimize.go:81-105)
		if err := andWith(scan.Filter); err != nil {
			return nil, 0, err
		}
	}
	i := 1
	for ; i < len(nodes) && nodes[i].Filter != nil; i++ {
		if err := andWith(nodes[i].Filter.Expr); err != nil {
			return nil, 0, err
		}
	}
	// the sqlparse pre-projection (visitor.go:62-130): plain columns, arithmetic, aliased arithmetic (computed keys)
	aliases := map[string]logicalplan.Expr{}
	if i+1 < len(nodes) && nodes[i].Projection != nil && nodes[i+1].Aggregation != nil {
		for _, e := range nodes[i].Projection.Exprs {
			if a, isAlias := e.(*logicalplan.AliasExpr); isAlias {
				aliases[a.Alias] = a.Expr
			}
		}
		i++
	}
	var kind C.int32_t
	var text []string
	switch {
	case i < len(nodes) && nodes[i].Aggregation != nil:
		kind = C.FGPU_PLAN_AGGREGATE
		for _, g := range nodes[i].Aggregation.GroupExprs {
			if c, isCol := g.(*logicalplan.Column); isCol {
				if e, aliased := aliases[c.ColumnName]; aliased {
					g = e // the library names the column after the expression; GPUScan renames it back (rename map)
				}
			}
			gi, err := p.add(g)
			if err != nil {
				return nil, 0, err
			}
			p.groups = append(p.groups, gi)
		}
		for _, a := range nodes[i].Aggregation.AggExprs { // AggFunc values are the wire values (expr.go:718-729)
			ai, err := p.add(a.Expr)
			if err != nil {
				return nil, 0, err
			}
			p.aggs = append(p.aggs, C.fgpu_agg{_func: C.int32_t(a.Func), expr: ai})
			text = append(text, a.Name())
		}
		i++
	case i+1 < len(nodes) && nodes[i].Projection != nil && nodes[i+1].Distinct != nil:
		kind = C.FGPU_PLAN_DISTINCT
		for _, g := range nodes[i+1].Distinct.Exprs {
			gi, err := p.add(g)
			if err != nil {
				return nil, 0, err
			}
			p.groups = append(p.groups, gi)
		}
		text = append(text, "distinct")
		i += 2
	default:
		return nil, 0, ErrGPUUnsupported
	}
	p.finish(scan.TableName, kind, filter)
	p.text = strings.Join(text, ",")
	ok = true
	return p, i, nil
}

// GPUScan implements ScanPhysicalPlan (physicalplan.go:32-35).
type GPUScan struct {
	gpu    *C.fgpu_ctx
	table  logicalplan.TableReader // View(): the read transaction (logicalplan.go:221-238)
	plan   *cPlan
	rename map[string]string // expression name -> alias of a computed group key
	next   PhysicalPlan
}

func (s *GPUScan) SetNext(p PhysicalPlan) { s.next = p }

func (s *GPUScan) Draw() *Diagram {
	return &Diagram{Details: "GPUScan (" + s.plan.text + ")", Child: s.next.Draw()}
}

func (s *GPUScan) Close() {
	s.plan.free()
	s.next.Close()
}

// Execute mirrors TableScan.Execute (physicalplan.go:114-166): take the read transaction, run the fused prefix on
// the GPU, push every result record into the next operator, then Finish it.  fgpu_query_prepare per Execute is
// cheap: equal plans share their compiled state inside the library.
func (s *GPUScan) Execute(ctx context.Context, _ memory.Allocator) error {
	return s.table.View(ctx, func(ctx context.Context, tx uint64) error {
		var q *C.fgpu_query
		if rc := C.fgpu_query_prepare(s.gpu, &s.plan.plan, &q); rc != 0 {
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
			rec, err := cdata.ImportCRecordBatch(&ca, &cs) // zero copy; the buffers go back to the library's pool on Release
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

// Patch 2, physicalplan.go:333 (first case of the plan visit in Build), when execOpts.gpu != nil:
//
//	case plan.TableScan != nil && execOpts.gpu != nil:
//		nodes := scanFirst(s)                       // the LogicalPlan chain, scan first
//		cp, used, err := encodeGPUPlan(nodes)
//		if err == nil {
//			table, terr := plan.TableScan.TableProvider.GetTable(plan.TableScan.TableName)
//			if terr != nil { return nil, terr }
//			gpuScan := &GPUScan{gpu: (*C.fgpu_ctx)(execOpts.gpu), table: table, plan: cp}
//			outputPlan.scan = gpuScan
//			prev = append(prev[:0], gpuScan)            // the operators after nodes[used-1] chain behind it
//			skip = used - 1                             // the visit skips the nodes the GPUScan consumed
//			break
//		}
//		if !errors.Is(err, ErrGPUUnsupported) { return nil, err }
//		fallthrough                                   // the existing TableScan case: Go operators
//
// An FGPU_ERR_UNSUPPORTED from fgpu_query_prepare / execute surfaces as ErrGPUUnsupported from Execute; the engine
// (query/engine.go:158-168) may then rebuild the plan without WithGPUEngine — the caller's choice.
