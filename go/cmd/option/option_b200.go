//go:build b200 && (badger || tikv)

// Compile-time storage selection (cf. cmd/option/option_badger.go:15-50, option_tikv.go):
//
//	go build --tags "badger b200"
//
// storageConfig below REPLACES the one of option_badger.go / option_tikv.go, whose build constraints a maintainer
// narrows to `badger && !b200` / `tikv && !b200` (one line each, INTEGRATION.md).  The durable engine is still built by
// the engine's own package; buildStorage wraps it in the HBM mirror (pkg/storage/b200).
package option

import (
	"fmt"
	"time"

	"github.com/spf13/pflag"

	"github.com/kubewharf/kubebrain/pkg/storage"
	"github.com/kubewharf/kubebrain/pkg/storage/b200"
)

type storageConfig struct {
	engine engineConfig // option_b200_badger.go / option_b200_tikv.go: the wrapped engine's own config
	b200.Config
}

func newStorageConfig() *storageConfig {
	s := &storageConfig{engine: newEngineConfig()}
	s.Device = 0
	s.ExpireTick = time.Second
	return s
}

func (s *storageConfig) addFlag(fs *pflag.FlagSet) {
	s.engine.addFlag(fs)
	fs.IntVar(&s.Device, "b200-device", s.Device, "CUDA device ordinal holding the HBM-resident snapshot")
	fs.StringVar(&s.SnapshotPath, "b200-snapshot", s.SnapshotPath,
		"file the HBM snapshot is dumped to on shutdown and restored from at start-up (empty: iterate the engine)")
	fs.DurationVar(&s.ExpireTick, "b200-expire-tick", s.ExpireTick, "period of the TTL sweep of the HBM mirror")
}

func (s *storageConfig) validate() error {
	if err := s.engine.validate(); err != nil {
		return err
	}
	if s.Device < 0 {
		return fmt.Errorf("b200-device must be >= 0, got %d", s.Device)
	}
	if s.ExpireTick <= 0 {
		return fmt.Errorf("b200-expire-tick must be positive, got %s", s.ExpireTick)
	}
	return nil
}

func (s *storageConfig) buildStorage() (storage.KvStorage, error) {
	inner, err := s.engine.build()
	if err != nil {
		return nil, err
	}
	return b200.NewKvStorage(inner, s.Config)
}
