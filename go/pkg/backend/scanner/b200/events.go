//go:build b200

package b200

import (
	proto "github.com/kubewharf/kubebrain-client/api/v2rpc"
)

// EventSlab is the revision-ordered slab kb_watch_match consumes (include/kb_b200.h kb_events): user keys back to
// back + offsets, Event.Revision per event, and the boundaries of the collector's <= 300-event batches.
//
// It replaces the hand-over at the end of collectStorageWriteEvents (pkg/backend/backend.go:208-270): where the
// reference copies the batch and pushes it into watchChan for WatcherHub.Stream to broadcast to every watcher
// (watcherhub.go:78-92, every watcher then filters it, watch.go:119-159), the collector appends the batch here; the hub
// goroutine takes the slab (several batches when it lags), runs ONE Engine.Match over it and hands every watcher only the
// events it is owed.  Batch boundaries are kept because filterByRevision strips only the LEADING events of a batch
// (watch.go:153-159).  The watch cache (b.watchCache.Add, backend.go:259) is fed by the collector exactly as before.
type EventSlab struct {
	Keys     []byte
	KeyOff   []uint64
	Rev      []uint64
	BatchOff []uint64
	Events   []*proto.Event // the messages themselves, index-aligned with Rev: what is delivered
}

func NewEventSlab() *EventSlab {
	return &EventSlab{KeyOff: []uint64{0}, BatchOff: []uint64{0}}
}

// AppendBatch adds one collector batch (evs[:cnt] of backend.go:263-266), in order.
func (s *EventSlab) AppendBatch(evs []*proto.Event) {
	if len(evs) == 0 {
		return
	}
	for _, e := range evs {
		s.Keys = append(s.Keys, e.Kv.Key...)
		s.KeyOff = append(s.KeyOff, uint64(len(s.Keys)))
		s.Rev = append(s.Rev, e.Revision)
		s.Events = append(s.Events, e)
	}
	s.BatchOff = append(s.BatchOff, uint64(len(s.Rev)))
}

func (s *EventSlab) Len() int { return len(s.Rev) }

// Reset empties the slab, keeping its capacity.
func (s *EventSlab) Reset() {
	s.Keys, s.KeyOff, s.Rev, s.BatchOff, s.Events = s.Keys[:0], s.KeyOff[:1], s.Rev[:0], s.BatchOff[:1], s.Events[:0]
}

// Deliveries is one watcher's share of a matched slab.
type Deliveries struct {
	Watcher uint32
	Events  []*proto.Event
}

// Fanout matches the slab against every registered watcher and returns, per watcher that is owed something, its events
// in stream order cut into the SAME per-batch messages the reference's processEvents would have sent (one message per
// collector batch with at least one surviving event, watch.go:128-150).
func (e *Engine) Fanout(s *EventSlab) (map[uint32][][]*proto.Event, error) {
	start, idx, err := e.Match(s.Keys, s.KeyOff, s.Rev, s.BatchOff)
	if err != nil || len(start) == 0 {
		return nil, err
	}
	// batch of every event: BatchOff is ascending, events are visited in ascending order per watcher
	out := make(map[uint32][][]*proto.Event)
	for w := 0; w+1 < len(start); w++ {
		lo, hi := start[w], start[w+1]
		if lo == hi {
			continue
		}
		var msgs [][]*proto.Event
		b := 0
		cur := -1
		for _, ei := range idx[lo:hi] {
			for uint64(ei) >= s.BatchOff[b+1] {
				b++
			}
			if b != cur {
				msgs = append(msgs, nil)
				cur = b
			}
			msgs[len(msgs)-1] = append(msgs[len(msgs)-1], s.Events[ei])
		}
		out[uint32(w)] = msgs
	}
	return out, nil
}

// SlowWatcherDropped is called by the hub when a watcher's channel is full (the select default branch of
// WatcherHub.Stream, watcherhub.go:84-89); the metric name is the reference's.
func SlowWatcherDropped(m Metrics) {
	if m != nil {
		_ = m.EmitCounter("drop.slow.watcher", 1)
	}
}
