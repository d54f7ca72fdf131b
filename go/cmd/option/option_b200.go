//go:build b200

// Compile-time storage selection, next to option_badger.go / option_tikv.go (cmd/option/option_badger.go:15-16):
//   go build --tags "badger b200"
// The durable engine stays badger; the B200 engine mirrors its snapshot in HBM and serves the scans.
package option

import (
	"github.com/spf13/pflag"
)

type b200Config struct {
	Device int
}

func (c *b200Config) addFlag(fs *pflag.FlagSet) {
	fs.IntVar(&c.Device, "b200-device", 0, "CUDA device ordinal holding the HBM-resident snapshot")
}
