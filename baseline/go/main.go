// Reference arm B0 (BASELINE.md §2): the UNMODIFIED polarsignals/frostdb Go engine on the same synthetic
// workload and query bench.py times, for a maintainer to run next to bench.py on a box that has a Go toolchain.
//
//	STATUS: UNRUN.  This image has no Go toolchain, no module cache and no network, so this file has never
//	been compiled; bench.py --impl reference times the C port of the same chain (oracle/) instead.
//
// Build (inside a checkout of github.com/polarsignals/frostdb, go 1.24):  go run ./baseline/go -rows 100000000
//
// Generator = bench_data.py: u(k,i) = splitmix64(S ^ k*GOLD ^ i), S = 0xF205DB; labels.l00..l15 with
// cardinalities 64, 256, 16, 32, 8, 128, 4, 64, 16, ... (NULL with probability 10 % for k >= 2);
// timestamp = T0 + i; value = u(200, i) mod 1000.  Rows go in through Table.InsertRecord in parts of 4 Mi rows
// and are compacted (EnsureCompaction) before the timed region, as bench.py's parts are.
// Query (headline): Filter(timestamp in the middle 50 %) -> Sum(value), Count(value) GROUP BY labels.l00, labels.l01.
package main

import (
	"context"
	"flag"
	"fmt"
	"runtime"
	"time"

	"github.com/apache/arrow-go/v18/arrow"
	"github.com/apache/arrow-go/v18/arrow/memory"

	"github.com/polarsignals/frostdb"
	"github.com/polarsignals/frostdb/query"
	"github.com/polarsignals/frostdb/query/logicalplan"
	"github.com/polarsignals/frostdb/samples"
)

const (
	seed     = uint64(0xF205DB)
	gold     = uint64(0x9E3779B97F4A7C15)
	t0       = int64(1_600_000_000_000)
	partRows = 4 * 1024 * 1024
)

func splitmix64(x uint64) uint64 {
	z := x + gold
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9
	z = (z ^ (z >> 27)) * 0x94D049BB133111EB
	return z ^ (z >> 31)
}

func u(k uint64, i uint64) uint64 { return splitmix64(seed ^ (k * gold) ^ i) }

func cardinalities(n int) []uint64 {
	cycle := []uint64{16, 32, 8, 128, 4, 64}
	out := make([]uint64, n)
	for k := range out {
		switch k {
		case 0:
			out[k] = 64
		case 1:
			out[k] = 256
		default:
			out[k] = cycle[(k-2)%len(cycle)]
		}
	}
	return out
}

func main() {
	rows := flag.Int("rows", 100_000_000, "rows of the table")
	labels := flag.Int("labels", 16, "dynamic label columns")
	steps := flag.Int("steps", 3, "timed executions")
	flag.Parse()
	ctx := context.Background()

	store, err := frostdb.New()
	must(err)
	defer store.Close()
	db, err := store.DB(ctx, "bench")
	must(err)
	table, err := db.Table("bench", frostdb.NewTableConfig(samples.SampleDefinition()))
	must(err)

	cards := cardinalities(*labels)
	for first := 0; first < *rows; first += partRows {
		n := min(partRows, *rows-first)
		part := make(samples.Samples, 0, n)
		for j := 0; j < n; j++ {
			i := uint64(first + j)
			ls := make(map[string]string, *labels)
			for k := 0; k < *labels; k++ {
				if k >= 2 && u(uint64(100+k), i)%100 < 10 {
					continue // NULL
				}
				ls[fmt.Sprintf("l%02d", k)] = fmt.Sprintf("v%06d", u(uint64(k), i)%cards[k])
			}
			part = append(part, samples.Sample{ExampleType: "cpu", Labels: ls, Timestamp: t0 + int64(i), Value: int64(u(200, i) % 1000)})
		}
		rec, err := part.ToRecord()
		must(err)
		_, err = table.InsertRecord(ctx, rec)
		must(err)
		rec.Release()
		must(table.EnsureCompaction())
	}

	lo, hi := t0+int64(*rows/4), t0+int64(3*(*rows)/4)
	engine := query.NewEngine(memory.DefaultAllocator, db.TableProvider())
	run := func() (groups int64) {
		must(engine.ScanTable("bench").
			Filter(logicalplan.And(logicalplan.Col("timestamp").GtEq(logicalplan.Literal(lo)), logicalplan.Col("timestamp").Lt(logicalplan.Literal(hi)))).
			Aggregate(
				[]*logicalplan.AggregationFunction{logicalplan.Sum(logicalplan.Col("value")), logicalplan.Count(logicalplan.Col("value"))},
				[]logicalplan.Expr{logicalplan.Col("labels.l00"), logicalplan.Col("labels.l01")},
			).Execute(ctx, func(_ context.Context, r arrow.Record) error {
			groups += r.NumRows()
			return nil
		}))
		return groups
	}
	run() // warm-up
	start := time.Now()
	var groups int64
	for s := 0; s < *steps; s++ {
		groups = run()
	}
	dt := time.Since(start).Seconds() / float64(*steps)
	fmt.Printf(`{"impl": "reference-go", "metric": "rows/sec scan+filter+hash-agg (100M-row Parca schema)", "value": %.0f, "unit": "rows/s", "cores": %d, "ms_per_step": %.3f, "groups": %d}`+"\n",
		float64(*rows)/dt, runtime.GOMAXPROCS(0), dt*1e3, groups)
}

func must(err error) {
	if err != nil {
		panic(err)
	}
}
